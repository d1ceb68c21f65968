/*
 * TEST INFRASTRUCTURE ONLY (oracle/): stand-in for libgfortran.so.3.
 *
 * The reference ships prebuilt f2py modules (fortran/waterlib.cpython-37m-x86_64-linux-gnu.so,
 * built with GCC 5.4) that list libgfortran.so.3 as NEEDED.  That runtime is not in this image and
 * there is no Fortran compiler, so the oracle loads the reference binary on top of this stub.  The
 * routines on the water-structure hot path (allnearneighbors_, nearneighbors_, reimage_,
 * tetracosang_, cosangle3_, generalhbonds_, lsidists_, histrr3b_, interfacewater_) never reach a
 * libgfortran entry point except through Fortran `stop` / `print`, so I/O entries are no-ops; the array
 * (un)packing helpers are implemented (watorient_ passes strided sections through them) and everything else
 * aborts loudly.
 */
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define WOL_NOOP(name) void name(void *a) { (void)a; }
#define WOL_FATAL(name) \
    void name(void) { fprintf(stderr, "oracle gfortran stub: %s reached\n", #name); abort(); }

WOL_NOOP(_gfortran_st_write)
WOL_NOOP(_gfortran_st_write_done)
void _gfortran_transfer_character_write(void *a, void *b, int c) { (void)a; (void)b; (void)c; }
void _gfortran_transfer_integer_write(void *a, void *b, int c) { (void)a; (void)b; (void)c; }
void _gfortran_transfer_real_write(void *a, void *b, int c) { (void)a; (void)b; (void)c; }

WOL_FATAL(_gfortran_stop_string)
WOL_FATAL(_gfortran_stop_numeric_f08)

/* Array descriptor of GCC 5's libgfortran ABI (the reference binary was built with GCC 5.4): strides in elements,
 * rank in the low 3 bits of dtype, element size in dtype >> 6.  internal_pack hands an explicit-shape dummy a
 * contiguous copy of a strided section (or the section itself when it already is contiguous); internal_unpack copies
 * the values back.  The caller frees the copy.  Reached from watorient_ (fortran/waterlib.f90:973-1011). */
typedef struct { ptrdiff_t stride, lbound, ubound; } wol_dim_t;
typedef struct { char *base_addr; size_t offset; ptrdiff_t dtype; wol_dim_t dim[7]; } wol_desc_t;

static ptrdiff_t wol_desc_walk(const wol_desc_t *d, char *packed, int to_packed) {
    const int rank = (int)(d->dtype & 7);
    const ptrdiff_t size = d->dtype >> 6;
    ptrdiff_t extent[7], idx[7], total = 1;
    for (int n = 0; n < rank; ++n) {
        extent[n] = d->dim[n].ubound - d->dim[n].lbound + 1;
        if (extent[n] <= 0) return 0;
        total *= extent[n];
        idx[n] = 0;
    }
    if (!packed) return total;
    for (ptrdiff_t k = 0; k < total; ++k) {
        ptrdiff_t off = 0;
        for (int n = 0; n < rank; ++n) off += idx[n] * d->dim[n].stride;
        if (to_packed) memcpy(packed + k * size, d->base_addr + off * size, (size_t)size);
        else memcpy(d->base_addr + off * size, packed + k * size, (size_t)size);
        for (int n = 0; n < rank; ++n) {
            if (++idx[n] < extent[n]) break;
            idx[n] = 0;
        }
    }
    return total;
}

void *_gfortran_internal_pack(wol_desc_t *d) {
    const int rank = (int)(d->dtype & 7);
    ptrdiff_t expect = 1;
    int contiguous = 1;
    for (int n = 0; n < rank; ++n) {
        const ptrdiff_t extent = d->dim[n].ubound - d->dim[n].lbound + 1;
        if (extent <= 0) return d->base_addr;
        if (d->dim[n].stride != expect) contiguous = 0;
        expect *= extent;
    }
    if (contiguous) return d->base_addr;
    char *packed = malloc((size_t)(expect * (d->dtype >> 6)));
    if (!packed) { fprintf(stderr, "oracle gfortran stub: out of memory in internal_pack\n"); abort(); }
    wol_desc_walk(d, packed, 1);
    return packed;
}

void _gfortran_internal_unpack(wol_desc_t *d, const void *src) {
    if (!src || src == (const void *)d->base_addr) return;
    wol_desc_walk(d, (char *)src, 0);
}

/* matmul for real(8), rank-2 x rank-1 (RadialDistPlane, fortran/waterlib.f90:278-292, is the only caller on a path the
 * oracle runs).  Same accumulation order as libgfortran 5's generic matmul_r8: dest(x) starts at zero and receives
 * a(x, n) * b(n) for n = 1, 2, ... in turn.  An unallocated result descriptor is allocated here (the caller frees it). */
void _gfortran_matmul_r8(wol_desc_t *ret, const wol_desc_t *a, const wol_desc_t *b, int try_blas, int blas_limit, void *gemm) {
    (void)try_blas; (void)blas_limit; (void)gemm;
    if ((a->dtype & 7) != 2 || (b->dtype & 7) != 1) {
        fprintf(stderr, "oracle gfortran stub: matmul_r8 only handles matrix x vector\n");
        abort();
    }
    const ptrdiff_t rows = a->dim[0].ubound - a->dim[0].lbound + 1, cols = a->dim[1].ubound - a->dim[1].lbound + 1;
    if (!ret->base_addr) {
        ret->dim[0].stride = 1;
        ret->dim[0].lbound = 0;
        ret->dim[0].ubound = rows - 1;
        ret->offset = 0;
        ret->base_addr = malloc((size_t)(rows > 0 ? rows : 1) * sizeof(double));
        if (!ret->base_addr) { fprintf(stderr, "oracle gfortran stub: out of memory in matmul_r8\n"); abort(); }
    }
    double *dest = (double *)ret->base_addr;
    const double *A = (const double *)a->base_addr, *B = (const double *)b->base_addr;
    const ptrdiff_t rs = ret->dim[0].stride ? ret->dim[0].stride : 1, as0 = a->dim[0].stride ? a->dim[0].stride : 1;
    const ptrdiff_t as1 = a->dim[1].stride, bs = b->dim[0].stride ? b->dim[0].stride : 1;
    for (ptrdiff_t x = 0; x < rows; ++x) dest[x * rs] = 0.0;
    for (ptrdiff_t n = 0; n < cols; ++n)
        for (ptrdiff_t x = 0; x < rows; ++x) dest[x * rs] += A[x * as0 + n * as1] * B[n * bs];
}
WOL_FATAL(_gfortran_random_r4)
WOL_FATAL(_gfortran_random_seed_i4)
WOL_FATAL(_gfortran_system_clock_4)
