/*
 * TEST INFRASTRUCTURE ONLY (oracle/): stand-in for libgfortran.so.3.
 *
 * The reference ships prebuilt f2py modules (fortran/waterlib.cpython-37m-x86_64-linux-gnu.so,
 * built with GCC 5.4) that list libgfortran.so.3 as NEEDED.  That runtime is not in this image and
 * there is no Fortran compiler, so the oracle loads the reference binary on top of this stub.  The
 * routines on the water-structure hot path (allnearneighbors_, nearneighbors_, reimage_,
 * tetracosang_, cosangle3_, generalhbonds_, lsidists_, histrr3b_, interfacewater_) never reach a
 * libgfortran entry point except through Fortran `stop` / `print`, so I/O entries are no-ops and
 * everything else aborts loudly.
 */
#include <stdio.h>
#include <stdlib.h>

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
WOL_FATAL(_gfortran_internal_pack)
WOL_FATAL(_gfortran_internal_unpack)
WOL_FATAL(_gfortran_matmul_r8)
WOL_FATAL(_gfortran_random_r4)
WOL_FATAL(_gfortran_random_seed_i4)
WOL_FATAL(_gfortran_system_clock_4)
