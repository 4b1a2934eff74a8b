// patch.cu -- the dynamic patcher boundary on device state.
//
// The reference's second way to build a graph is at run time, over the TAG_U32 RPC tree
// of stm32f103/mod_bpmodular.c:
//     class/<c>/apply(in...)            -> allocate an instance of processor class c whose
//                                          inputs read the .out of existing nodes (:85-117,
//                                          :261-276); the reply is the node index
//     inst/<node>/state|param/<k>/get|set  (:152-189)
//     patch/reset, patch/tick           (:245-256)
// with the field names coming from the proc_meta tables (generic/cproc.h:107-122).
// Here the node list is the ANF node table of CPROC_CUDA_GRAPH, instances are N-wide
// batches, state lives in the SoA rows on the device, and `tick` renders F ticks with
// the NVRTC-compiled kernel of the current table.  Nodes can only be appended
// (balloci has no free either: "Individual node deletion is not supported", :243), so a
// rebuilt batch keeps every existing state row.
#include "common.cuh"
#include <string.h>

namespace {
// Patcher classes: the node kinds of the graph (field names from k_cproc_kinds: for_acc_state / for_edge_state of
// cproc.h:134-148, glide / pdm of cproc_cuda.h, the extension processors of include/cproc_ext.h) plus `input`, an external
// stream as a source node (the role gpin plays on the microcontroller, hw_cproc_stm32f103.h:8-14).  Class numbers 0..4 are
// those of ABI version 2.
enum { CLS_ACC = 0, CLS_EDGE = 1, CLS_GLIDE = 2, CLS_INPUT = 3, CLS_PDM = 4, CLS_PHASOR_F = 5, CLS_SVF = 6, CLS_ENV = 7, CLS_ONEPOLE = 8, CLS_GAIN = 9, CLS_ASFLOAT = 10, CLS_GLIDE_F = 11, CLS_MUL = 12 };
const uint32_t k_n_classes = 13;
const int k_class_kind[k_n_classes] = {CPROC_CUDA_NODE_ACC, CPROC_CUDA_NODE_EDGE, CPROC_CUDA_NODE_GLIDE, -1, CPROC_CUDA_NODE_PDM, CPROC_CUDA_NODE_PHASOR_F,
                                       CPROC_CUDA_NODE_SVF, CPROC_CUDA_NODE_ENV, CPROC_CUDA_NODE_ONEPOLE, CPROC_CUDA_NODE_GAIN, CPROC_CUDA_NODE_ASFLOAT,
                                       CPROC_CUDA_NODE_GLIDE_F, CPROC_CUDA_NODE_MUL};
const cproc_kind_meta k_input_class = {"input", 0, {nullptr}, 0, 0, {nullptr}, 0, 0, {nullptr}, 0, 1, {"index"}};
const cproc_kind_meta &class_meta(uint32_t cls) { return k_class_kind[cls] < 0 ? k_input_class : k_cproc_kinds[k_class_kind[cls]]; }
}  // namespace

struct cproc_cuda_patch {
    cproc_cuda_ctx *ctx = nullptr;
    uint64_t n = 0;
    uint32_t n_inputs = 0;
    struct PNode { uint32_t cls; int32_t table; uint32_t input; };   // table: row in `rows` (or -1 for an input node)
    std::vector<PNode> nodes;
    std::vector<cproc_cuda_node> rows;
    std::vector<uint32_t> off;          // state word offset per table row
    std::vector<uint32_t> poff;         // param word offset per table row
    int32_t out_node = -1;              // patch node index
    cproc_cuda_batch *batch = nullptr;
    bool dirty = true;
    uint32_t layout = CPROC_CUDA_PLANAR;
};

extern "C" {

int cproc_cuda_patch_class_count(void) { return (int)k_n_classes; }
const char *cproc_cuda_patch_class_name(uint32_t cls) { return cls < k_n_classes ? class_meta(cls).name : nullptr; }

// kind: 0 param, 1 state (PARAM / STATE of mod_bpmodular.c:126-127), 2 input, 3 config.  Returns the
// number of fields (or -1); name k through `name` when it is non-NULL.
int cproc_cuda_patch_class_field(uint32_t cls, uint32_t kind, uint32_t k, const char **name) {
    if (cls >= k_n_classes) return -1;
    const cproc_kind_meta &m = class_meta(cls);
    uint32_t n = 0; const char *const *names = nullptr;
    switch (kind) {
    case 0: n = m.n_param; names = m.param; break;
    case 1: n = m.n_state; names = m.state; break;
    case 2: n = m.n_input; names = m.input; break;
    case 3: n = m.n_config; names = m.config; break;
    default: return -1;
    }
    if (name) *name = (names && k < n) ? names[k] : nullptr;
    return (int)n;
}

int cproc_cuda_patch_open(cproc_cuda_ctx *ctx, uint64_t n_instances, uint32_t n_inputs, uint32_t layout, cproc_cuda_patch **out) {
    if (!ctx || !out) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "patch_open: NULL argument");
    *out = nullptr;
    if (n_instances == 0) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "patch_open: n_instances is 0");
    if (layout != CPROC_CUDA_PLANAR && layout != CPROC_CUDA_INTERLEAVED) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "patch_open: layout must be PLANAR or INTERLEAVED");
    cproc_cuda_patch *p = new cproc_cuda_patch();
    p->ctx = ctx; p->n = n_instances; p->n_inputs = n_inputs; p->layout = layout;
    *out = p;
    return 0;
}

// patch/reset (mod_bpmodular.c:245-249: balloci_clear)
int cproc_cuda_patch_reset(cproc_cuda_patch *p) {
    if (!p) return cproc_set_err(nullptr, CPROC_CUDA_EINVAL, "patch_reset: patch is NULL");
    if (p->batch) { cproc_cuda_free(p->batch); p->batch = nullptr; }
    p->nodes.clear(); p->rows.clear(); p->off.clear(); p->poff.clear(); p->out_node = -1; p->dirty = true;
    return 0;
}

int cproc_cuda_patch_close(cproc_cuda_patch *p) {
    if (!p) return 0;
    cproc_cuda_patch_reset(p);
    delete p;
    return 0;
}

int cproc_cuda_patch_node_count(const cproc_cuda_patch *p) { return p ? (int)p->nodes.size() : 0; }

// class/<cls>/apply: returns the new node index (>= 0) or a negative error ("bad_node" /
// "bad_ref" / "alloc_fail" of the reference are EINVAL / EINVAL / ENOMEM here).
int cproc_cuda_patch_apply(cproc_cuda_patch *p, uint32_t cls, const uint32_t *in_nodes, uint32_t n_in, uint32_t config) {
    if (!p) return cproc_set_err(nullptr, CPROC_CUDA_EINVAL, "patch_apply: patch is NULL");
    cproc_cuda_ctx *ctx = p->ctx;
    if (cls >= k_n_classes) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "patch_apply: bad_ref: class %u", cls);
    const cproc_kind_meta &m = class_meta(cls);
    // the extension processors' inputs may stay unconnected (an omitted member of the C initialiser reads 0): n_in <= n_input
    const bool ext = k_class_kind[cls] > CPROC_CUDA_NODE_PDM;
    if (ext ? n_in > m.n_input : n_in != m.n_input) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "patch_apply: bad_ref: %s takes %u inputs, got %u", m.name, m.n_input, n_in);   // :263
    if (n_in && !in_nodes) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "patch_apply: inputs are NULL");
    cproc_cuda_patch::PNode pn{cls, -1, 0};
    if (cls == CLS_INPUT) {
        if (config >= p->n_inputs) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "patch_apply: bad_ref: input stream %u of %u", config, p->n_inputs);
        pn.input = config;
    } else {
        if (p->rows.size() >= CPROC_CUDA_GRAPH_MAX_NODES) return cproc_set_err(ctx, CPROC_CUDA_ENOMEM, "patch_apply: alloc_fail: %d nodes", CPROC_CUDA_GRAPH_MAX_NODES);   // :94-97
        cproc_cuda_node row;
        row.cond_mask = 0xFFFFFFFFu;                        // tick() runs every instance (:71-77)
        row.src = CPROC_CUDA_SRC_ZERO; row.src2 = 0;
        for (uint32_t j = 0; j < n_in; ++j) {
            if (in_nodes[j] >= p->nodes.size()) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "patch_apply: bad_node: %u", in_nodes[j]);                                    // :106-110
            const cproc_cuda_patch::PNode &src = p->nodes[in_nodes[j]];
            const int32_t sv = src.table >= 0 ? src.table : -(int32_t)src.input - 1;
            if (j == 0) row.src = sv; else row.src2 = sv;
        }
        if (cls == CLS_PDM) {
            if ((config & 7u) < 1 || (config & 7u) > 4 || (config >> 3) > 31) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "patch_apply: pdm config must be order 1..4 | out_shift << 3");
            row.type = CPROC_CUDA_NODE_PDM | (config << 8);
        } else if (cls == CLS_GLIDE || cls == CLS_GLIDE_F) {
            if (config < 1 || config > 24) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "patch_apply: glide div_log must be 1..24");
            row.type = (uint32_t)k_class_kind[cls] | (config << 8);
        } else {
            if (config) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "patch_apply: %s takes no config word", m.name);
            row.type = (uint32_t)k_class_kind[cls];
        }
        p->rows.push_back(row);
        if (const char *why = cproc_node_check(p->rows.data(), (uint32_t)p->rows.size() - 1, p->n_inputs)) {
            p->rows.pop_back();
            return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "patch_apply: bad_ref: %s: %s", m.name, why);
        }
        const size_t r = p->rows.size() - 1;
        p->off.push_back(r ? p->off[r - 1] + cproc_node_words(p->rows[r - 1].type) : 0);
        p->poff.push_back(r ? p->poff[r - 1] + cproc_node_param_words(p->rows[r - 1].type) : 0);
        if (p->poff[r] + cproc_node_param_words(row.type) > CPROC_CUDA_GRAPH_MAX_PARAM_WORDS) {
            p->rows.pop_back(); p->off.pop_back(); p->poff.pop_back();
            return cproc_set_err(ctx, CPROC_CUDA_ENOMEM, "patch_apply: alloc_fail: param record exceeds %d words", CPROC_CUDA_GRAPH_MAX_PARAM_WORDS);
        }
        pn.table = (int32_t)r;
        p->dirty = true;
    }
    p->nodes.push_back(pn);
    return (int)p->nodes.size() - 1;
}

// The node whose .out the tick renders into io->out (the reference's graphs end in gpout / cproc_output).
int cproc_cuda_patch_output(cproc_cuda_patch *p, uint32_t node) {
    if (!p) return cproc_set_err(nullptr, CPROC_CUDA_EINVAL, "patch_output: patch is NULL");
    if (node >= p->nodes.size() || p->nodes[node].table < 0) return cproc_set_err(p->ctx, CPROC_CUDA_EINVAL, "patch_output: bad_ref: node %u has no state", node);
    if (p->out_node != (int32_t)node) { p->out_node = (int32_t)node; p->dirty = true; }
    return 0;
}

static int patch_build(cproc_cuda_patch *p) {
    cproc_cuda_ctx *ctx = p->ctx;
    if (!p->dirty && p->batch) return 0;
    if (p->rows.empty()) return cproc_set_err(ctx, CPROC_CUDA_ESTATE, "patch: no processor instance yet");
    cproc_cuda_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.proc = CPROC_CUDA_GRAPH; cfg.layout = p->layout;
    cfg.nodes = p->rows.data(); cfg.n_nodes = (uint32_t)p->rows.size(); cfg.n_inputs = p->n_inputs;
    // no output selected yet (param / state access while the patch is being built): the last instance, like a graph that ends in its sink
    cfg.out_node = p->out_node >= 0 ? (uint32_t)p->nodes[p->out_node].table : (uint32_t)p->rows.size() - 1; cfg.n_outputs = 1; cfg.out_nodes = nullptr;
    cproc_cuda_batch *nb = nullptr;
    int rc = cproc_cuda_alloc(ctx, &cfg, p->n, &nb);
    if (rc) return rc;
    if (p->batch) {                                          // append-only: the old rows are a prefix of the new ones
        const size_t bytes = sizeof(uint32_t) * p->batch->state_words * p->batch->npad;
        rc = cproc_check(ctx, cudaMemcpyAsync(nb->d_state, p->batch->d_state, bytes, cudaMemcpyDeviceToDevice, ctx->stream), "patch: carry state");
        if (!rc && p->batch->param_words)
            rc = cproc_check(ctx, cudaMemcpyAsync(nb->d_param, p->batch->d_param, sizeof(uint32_t) * p->batch->param_words * p->batch->npad, cudaMemcpyDeviceToDevice, ctx->stream), "patch: carry params");
        if (rc) { cproc_cuda_free(nb); return rc; }
        cproc_cuda_free(p->batch);
    }
    p->batch = nb; p->dirty = false;
    return 0;
}

// patch/tick, F ticks at once (mod_bpmodular.c:251-255 runs one): io->in external streams, io->out the output node
int cproc_cuda_patch_tick(cproc_cuda_patch *p, uint64_t n_frames, const cproc_cuda_io *io, int device_buffers) {
    if (!p || !io) return cproc_set_err(p ? p->ctx : nullptr, CPROC_CUDA_EINVAL, "patch_tick: NULL argument");
    if (p->out_node < 0) return cproc_set_err(p->ctx, CPROC_CUDA_ESTATE, "patch: no output node selected (cproc_cuda_patch_output)");
    int rc = patch_build(p);
    if (rc) return rc;
    return device_buffers ? cproc_cuda_run_dev(p->batch, n_frames, io) : cproc_cuda_run(p->batch, n_frames, io);
}

static int patch_word(cproc_cuda_patch *p, uint32_t node, uint32_t kind, uint32_t field, uint64_t instance, uint32_t **dev) {
    cproc_cuda_ctx *ctx = p->ctx;
    if (node >= p->nodes.size()) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "patch: bad_ref: node %u", node);
    const cproc_cuda_patch::PNode &pn = p->nodes[node];
    const cproc_kind_meta &m = class_meta(pn.cls);
    if (kind > 1) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "patch: bad_ref: kind %u (0 param, 1 state)", kind);
    if (kind == 0 && m.n_param == 0) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "patch: bad_ref: %s has no param fields", m.name);   // acc / edge: params are not stored in the reference either (:166, :184)
    const uint32_t n_fields = pn.table < 0 ? 0 : kind == 0 ? m.n_param : cproc_node_words(p->rows[pn.table].type);   // pdm: 1 + order of this instance
    if (field >= n_fields) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "patch: bad_ref: %s %s field %u", m.name, kind ? "state" : "param", field);
    if (instance >= p->n) return cproc_set_err(ctx, CPROC_CUDA_EINVAL, "patch: bad_ref: instance %llu", (unsigned long long)instance);
    int rc = patch_build(p);
    if (rc) return rc;
    *dev = kind == 0 ? p->batch->d_param + (uint64_t)(p->poff[pn.table] + field) * p->batch->npad + instance
                     : p->batch->d_state + (uint64_t)(p->off[pn.table] + field) * p->batch->npad + instance;
    return 0;
}

// inst/<node>/state/<field>/get  (mod_bpmodular.c:152-170)
int cproc_cuda_patch_get(cproc_cuda_patch *p, uint32_t node, uint32_t kind, uint32_t field, uint64_t instance, uint32_t *value) {
    if (!p || !value) return cproc_set_err(p ? p->ctx : nullptr, CPROC_CUDA_EINVAL, "patch_get: NULL argument");
    uint32_t *d = nullptr;
    int rc = patch_word(p, node, kind, field, instance, &d);
    if (rc) return rc;
    CK(p->ctx, cudaMemcpyAsync(value, d, 4, cudaMemcpyDeviceToHost, p->ctx->stream));
    CK(p->ctx, cudaStreamSynchronize(p->ctx->stream));
    return 0;
}

// inst/<node>/state/<field>/set  (:172-189)
int cproc_cuda_patch_set(cproc_cuda_patch *p, uint32_t node, uint32_t kind, uint32_t field, uint64_t instance, uint32_t value) {
    if (!p) return cproc_set_err(nullptr, CPROC_CUDA_EINVAL, "patch_set: patch is NULL");
    uint32_t *d = nullptr;
    int rc = patch_word(p, node, kind, field, instance, &d);
    if (rc) return rc;
    CK(p->ctx, cudaMemcpyAsync(d, &value, 4, cudaMemcpyHostToDevice, p->ctx->stream));
    CK(p->ctx, cudaStreamSynchronize(p->ctx->stream));
    return 0;
}

// the batch behind the patch (state checkpoint: cproc_cuda_download_state / upload_state on it)
cproc_cuda_batch *cproc_cuda_patch_batch(cproc_cuda_patch *p) {
    if (!p || patch_build(p)) return nullptr;
    return p->batch;
}

}  // extern "C"
