/* Minimal Pd typedefs so linux/synth_tools.c:78-100 compiles stand-alone
 * (m_pd.h is not in this image).  t_object is opaque here: the DSP code
 * never touches x_obj. */
#include <stdint.h>
#include <stddef.h>
typedef float t_float;
typedef intptr_t t_int;
typedef struct { void *ob_pd; void *ob_binbuf; void *ob_inlet; char ob_type; } t_object;
typedef struct _class t_class;
