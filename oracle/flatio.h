/* flatio.h -- TEST INFRASTRUCTURE (oracle/). Tiny named-array container ("AWF1")
 * used to move flattened thread-sampling problems and reference outputs between
 * the reference-linked drivers (ref_dump / ref_bench), the C oracle and Python.
 *
 * File layout:  "AWF1"  then records until EOF:
 *   u32 name_len | name bytes | u8 dtype (0=i32 1=f64 2=u8 3=i64) | u32 ndim |
 *   u64 dims[ndim] | raw little-endian data
 *
 * The Python twin of this reader/writer is argweaver_b200/flatfile.py.
 */
#ifndef AWB_ORACLE_FLATIO_H
#define AWB_ORACLE_FLATIO_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

enum { AWF_I32 = 0, AWF_F64 = 1, AWF_U8 = 2, AWF_I64 = 3 };

static inline size_t awf_dtype_size(int dtype)
{
    switch (dtype) {
    case AWF_I32: return 4;
    case AWF_F64: return 8;
    case AWF_U8:  return 1;
    case AWF_I64: return 8;
    }
    return 0;
}

/* ---------------------------------------------------------------- writer */

static inline FILE *awf_create(const char *filename)
{
    FILE *f = fopen(filename, "wb");
    if (f) fwrite("AWF1", 1, 4, f);
    return f;
}

static inline void awf_write(FILE *f, const char *name, int dtype, int ndim,
                             const uint64_t *dims, const void *data)
{
    uint32_t nl = (uint32_t) strlen(name);
    uint8_t dt = (uint8_t) dtype;
    uint32_t nd = (uint32_t) ndim;
    uint64_t n = 1;
    for (int i = 0; i < ndim; i++) n *= dims[i];
    fwrite(&nl, 4, 1, f);
    fwrite(name, 1, nl, f);
    fwrite(&dt, 1, 1, f);
    fwrite(&nd, 4, 1, f);
    fwrite(dims, 8, ndim, f);
    if (n) fwrite(data, awf_dtype_size(dtype), n, f);
}

static inline void awf_write1(FILE *f, const char *name, int dtype,
                              uint64_t n, const void *data)
{
    awf_write(f, name, dtype, 1, &n, data);
}

static inline void awf_write2(FILE *f, const char *name, int dtype,
                              uint64_t n0, uint64_t n1, const void *data)
{
    uint64_t dims[2] = { n0, n1 };
    awf_write(f, name, dtype, 2, dims, data);
}

static inline void awf_write_int(FILE *f, const char *name, int v)
{
    int32_t x = v;
    awf_write1(f, name, AWF_I32, 1, &x);
}

static inline void awf_write_double(FILE *f, const char *name, double v)
{
    awf_write1(f, name, AWF_F64, 1, &v);
}

/* ---------------------------------------------------------------- reader */

typedef struct {
    char name[64];
    int dtype;
    int ndim;
    uint64_t dims[4];
    uint64_t count;
    void *data;
} awf_array;

typedef struct {
    int narrays;
    awf_array *arrays;
} awf_file;

static inline awf_file *awf_read(const char *filename)
{
    FILE *f = fopen(filename, "rb");
    char magic[4];
    if (!f) return NULL;
    if (fread(magic, 1, 4, f) != 4 || memcmp(magic, "AWF1", 4) != 0) {
        fclose(f);
        return NULL;
    }
    awf_file *af = (awf_file *) calloc(1, sizeof(awf_file));
    int cap = 64;
    af->arrays = (awf_array *) calloc(cap, sizeof(awf_array));
    for (;;) {
        uint32_t nl, nd;
        uint8_t dt;
        if (fread(&nl, 4, 1, f) != 1) break;
        if (af->narrays == cap) {
            cap *= 2;
            af->arrays = (awf_array *) realloc(af->arrays,
                                               cap * sizeof(awf_array));
        }
        awf_array *a = &af->arrays[af->narrays];
        memset(a, 0, sizeof(*a));
        if (nl >= sizeof(a->name)) { fclose(f); return NULL; }
        if (fread(a->name, 1, nl, f) != nl) break;
        a->name[nl] = 0;
        if (fread(&dt, 1, 1, f) != 1) break;
        if (fread(&nd, 4, 1, f) != 1) break;
        a->dtype = dt;
        a->ndim = (int) nd;
        a->count = 1;
        for (uint32_t i = 0; i < nd; i++) {
            if (fread(&a->dims[i], 8, 1, f) != 1) break;
            a->count *= a->dims[i];
        }
        size_t bytes = a->count * awf_dtype_size(a->dtype);
        a->data = malloc(bytes ? bytes : 1);
        if (bytes && fread(a->data, 1, bytes, f) != bytes) break;
        af->narrays++;
    }
    fclose(f);
    return af;
}

static inline awf_array *awf_find(awf_file *af, const char *name)
{
    for (int i = 0; i < af->narrays; i++)
        if (strcmp(af->arrays[i].name, name) == 0)
            return &af->arrays[i];
    return NULL;
}

static inline awf_array *awf_need(awf_file *af, const char *name)
{
    awf_array *a = awf_find(af, name);
    if (!a) {
        fprintf(stderr, "awf: missing array '%s'\n", name);
        exit(2);
    }
    return a;
}

static inline int awf_int(awf_file *af, const char *name)
{
    return ((int32_t *) awf_need(af, name)->data)[0];
}

static inline double awf_double(awf_file *af, const char *name)
{
    return ((double *) awf_need(af, name)->data)[0];
}

static inline void awf_free(awf_file *af)
{
    if (!af) return;
    for (int i = 0; i < af->narrays; i++) free(af->arrays[i].data);
    free(af->arrays);
    free(af);
}

#endif /* AWB_ORACLE_FLATIO_H */
