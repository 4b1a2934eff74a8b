/* Stand-in for uc_tools metastruct.h (github:zwizwa/uc_tools, not vendored in
 * the reference tree; generic/cproc.h:48 includes it).  Only what cproc.h
 * needs to compile: a struct typedef from an X-macro field list, and the
 * reflection record types that struct proc_meta embeds (cproc.h:107-110,
 * used by stm32f103/mod_bpmodular.c:65).  Written for this repo; used only
 * to build oracle/_ref from the unmodified reference headers. */
#ifndef METASTRUCT_H
#define METASTRUCT_H
#include <stdint.h>
#define METASTRUCT_FIELD_DECL(type, name) type name;
#define METASTRUCT_CONST_FIELD_DECL(type, name) const type name;
#define STRUCT_DEF(name, for_fields) typedef struct { for_fields(METASTRUCT_FIELD_DECL) } name
#define STRUCT_CONST_DEF(name, for_fields) typedef struct { for_fields(METASTRUCT_CONST_FIELD_DECL) } name
struct metastruct_field { const char *type; const char *name; };
struct metastruct_struct { uint32_t nb_fields; const struct metastruct_field *fields; };
#define METASTRUCT_FIELD_META(type, name) { #type, #name },
#define METASTRUCT_DEF(name, for_fields) \
    static const struct metastruct_field name##_fields[] = { for_fields(METASTRUCT_FIELD_META) {0, 0} }
#define METASTRUCT_STRUCT(name) \
    { (sizeof(name##_fields) / sizeof(name##_fields[0])) - 1, name##_fields }
#endif
