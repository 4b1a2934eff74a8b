/* A scripted stand-in for the part of m_pd.h that synth_tools_b200/host/pd/square_grain_b200~.c uses
 * (this image has no Pd).  Test infrastructure: see fakepd.c. */
#ifndef FAKE_M_PD_H
#define FAKE_M_PD_H
#include <stdint.h>
typedef float t_float;
typedef float t_floatarg;
typedef intptr_t t_int;
typedef struct fake_symbol { const char *s_name; } t_symbol;
typedef struct fake_class t_class;
typedef t_class *t_pd;
typedef struct fake_object { t_pd ob_pd; void *pad[3]; } t_object;
typedef struct fake_signal { int s_n; t_float *s_vec; } t_signal;
typedef struct fake_inlet t_inlet;
typedef struct fake_outlet t_outlet;
typedef void (*t_method)(void);
typedef void *(*t_newmethod)(void);
typedef t_int *(*t_perfroutine)(t_int *w);
typedef enum { A_NULL = 0, A_FLOAT, A_SYMBOL, A_POINTER, A_SEMI, A_COMMA, A_DEFFLOAT, A_DEFSYM, A_DOLLAR, A_DOLLSYM, A_GIMME, A_CANT } t_atomtype;
#define CLASS_DEFAULT 0
#define A_DEFSYMBOL A_DEFSYM
#define CLASS_MAINSIGNALIN(c, type, field) class_domainsignalin(c, (int)((char *)&((type *)0)->field - (char *)0))
t_symbol *gensym(const char *s);
t_class *class_new(t_symbol *name, t_newmethod newmethod, t_method freemethod, size_t size, int flags, t_atomtype arg1, ...);
void class_addmethod(t_class *c, t_method fn, t_symbol *sel, t_atomtype arg1, ...);
void class_domainsignalin(t_class *c, int onset);
t_pd *pd_new(t_class *cls);
t_inlet *inlet_new(t_object *owner, t_pd *dest, t_symbol *s1, t_symbol *s2);
t_outlet *outlet_new(t_object *owner, t_symbol *s);
void dsp_add(t_perfroutine f, int n, ...);
void post(const char *fmt, ...);
void class_addfloat(t_class *c, t_method fn);
void outlet_float(t_outlet *x, t_float f);
#endif
