"""synth_tools_b200 -- B200-native batched renderer for the synth_tools
per-sample DSP hot path (cproc graphs, PDM modulators, phasor voice bank,
square_grain).  The product is libcproc_cuda.so behind include/cproc_cuda.h;
this package is the thin Python host binding over that C-ABI.

Importing the package loads the CUDA library and fails loudly when it has not
been built: there is no CPU fallback.
"""
from . import abi  # noqa: F401  (raises ImportError if libcproc_cuda.so is missing)
from .abi import (Batch, Bus, Context, Patch, CprocCudaError, GRAPH, INTERLEAVED, MIX_SAW, MIX_SQUARE, NODE_ACC, NODE_EDGE, NODE_GLIDE, NODE_PDM,
                  ONEPOLE, WORD_CLOCK, PDM, PDM_V1, PDM_V2, PLANAR, PWM, SQUARE_GRAIN, SQUARE_GRAIN_MIX, TILED, VOICE_BANK,
                  XVOICE, XVOICE_SCAN, XVOICE_SEQ, graph_parse, graph_parse_ex, graph_parse_outputs, node_glide, node_pdm,
                  NODE_PHASOR_F, NODE_SVF, NODE_ENV, NODE_ONEPOLE, NODE_GAIN, NODE_ASFLOAT, NODE_GLIDE_F, NODE_MUL, node_glide_f, SRC_ZERO)

__all__ = ["abi", "Batch", "Bus", "Context", "Patch", "CprocCudaError", "GRAPH", "PDM", "PDM_V1", "PDM_V2", "PWM", "VOICE_BANK",
           "SQUARE_GRAIN", "SQUARE_GRAIN_MIX", "XVOICE", "ONEPOLE", "WORD_CLOCK", "NODE_ACC", "NODE_EDGE", "NODE_GLIDE", "NODE_PDM", "node_glide", "node_pdm", "graph_parse", "graph_parse_ex", "graph_parse_outputs", "NODE_PHASOR_F", "NODE_SVF", "NODE_ENV", "NODE_ONEPOLE", "NODE_GAIN", "NODE_ASFLOAT", "NODE_GLIDE_F", "NODE_MUL", "node_glide_f", "SRC_ZERO", "MIX_SAW",
           "MIX_SQUARE", "XVOICE_SEQ", "XVOICE_SCAN", "PLANAR", "INTERLEAVED", "TILED"]
