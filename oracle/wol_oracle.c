/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the water-structure hot path.
 *
 * A plain-C, fp64 restatement of what the reference computes per frame, written so that every
 * floating-point operation happens in the order the reference performs it (no FMA contraction: the
 * reference's prebuilt Fortran is SSE2-only; build with -ffp-contract=off).  It exists so that the
 * CUDA path can be checked at sizes the reference's dense N x N neighbour matrices cannot reach
 * (fortran/waterlib.f90:836 is 4 GiB at N=32768).  It is itself pinned against the reference's
 * compiled Fortran + unmodified Python (oracle/ref_fortran.py) by tests/test_oracle_vs_reference.py
 * and against the .npz fixtures under tests/golden.
 *
 * The product (waterorderlib_b200/) never links, loads or calls this file.
 *
 * Reference anchors (relative to /root/reference):
 *   min-image idiom            fortran/waterlib.f90:41-44
 *   cutoff test                fortran/waterlib.f90:733-741, :851-859
 *   reimage                    fortran/waterlib.f90:32-47
 *   tetraCosAng / CosAngle3    fortran/waterlib.f90:867-895, :683-703
 *   generalHbonds / AngBetween fortran/waterlib.f90:1156-1210, :954-965
 *   getCosAngs                 structureLibs/water_properties.py:210-250
 *   getOrderParamq             structureLibs/water_properties.py:344-391
 *   tetrahedralMetrics         structureLibs/water_properties.py:314-342 (np.histogram uniform bins)
 *   shell selection            structureLibs/orderParam_lib.py:495-498
 *
 * All arrays are row-major: positions (n,3) double.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define WOL_PI 3.1415926535897931
static const double kPi = WOL_PI;
static const double kTwoPi = WOL_PI * 2.0;
static const double kDegPerRad = 180.0 / WOL_PI;

typedef struct {
    double L[3], iL[3];
} box_t;

static void box_init(box_t *b, const double *boxl) {
    for (int d = 0; d < 3; ++d) {
        b->L[d] = boxl[d];
        /* iBoxL = merge(1.d0/BoxL, 0.d0, BoxL >= 0.d0)  (waterlib.f90:41) */
        b->iL[d] = (boxl[d] >= 0.0) ? 1.0 / boxl[d] : 0.0;
    }
}

/* distvec = p - r ; distvec = distvec - BoxL * anint(distvec * iBoxL)  (waterlib.f90:43-44) */
static inline void min_image(const box_t *b, const double *p, const double *r, double *d) {
    for (int k = 0; k < 3; ++k) {
        double t = p[k] - r[k];
        double s = t * b->iL[k];
        d[k] = t - b->L[k] * round(s);
    }
}

static inline double sumsq(const double *v) { return (v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]; }

/* CosAngle3(Pos1, Pos2, Pos3) in degrees (waterlib.f90:683-703), Pos2 is the vertex */
static double cos_angle3(const double *p1, const double *p2, const double *p3) {
    if ((p1[0] == p2[0] && p1[1] == p2[1] && p1[2] == p2[2]) ||
        (p2[0] == p3[0] && p2[1] == p3[1] && p2[2] == p3[2]))
        return 0.0;
    double v21[3], v23[3];
    for (int k = 0; k < 3; ++k) {
        v21[k] = p1[k] - p2[k];
        v23[k] = p3[k] - p2[k];
    }
    double norm = sqrt(sumsq(v21) * sumsq(v23));
    double dot = (v21[0] * v23[0] + v21[1] * v23[1]) + v21[2] * v23[2];
    double phi = fmin(1.0, fmax(-1.0, dot / norm));
    phi = acos(phi);
    double a = fmod(phi + kPi, kTwoPi) - kPi;
    if (a < -kPi) a += kTwoPi;
    return a * kDegPerRad;
}

/* AngBetween(Vec1, Vec2) in degrees for normalised vectors (waterlib.f90:954-965) */
static double ang_between(const double *v1, const double *v2) {
    double dot = (v1[0] * v2[0] + v1[1] * v2[1]) + v1[2] * v2[2];
    double phi = acos(fmin(1.0, fmax(-1.0, dot)));
    double a = fmod(phi + kPi, kTwoPi) - kPi;
    if (a < -kPi) a += kTwoPi;
    return a * kDegPerRad;
}

/* ------------------------------------------------------------------------------------------ */
/* Cell list over Pos (periodic axes only; used when every axis has >= 3 cells of edge >= rc).  */

typedef struct {
    int nc[3];
    int *start; /* ncell + 1 */
    int *items; /* n, atom indices grouped by cell, ascending inside a cell */
    int ok;
} cells_t;

static int cell_of(const box_t *b, const int *nc, const double *p, int *c3) {
    for (int k = 0; k < 3; ++k) {
        double f = p[k] * b->iL[k];
        f -= floor(f);
        int c = (int)(f * nc[k]);
        if (c >= nc[k]) c = nc[k] - 1;
        if (c < 0) c = 0;
        c3[k] = c;
    }
    return (c3[2] * nc[1] + c3[1]) * nc[0] + c3[0];
}

static void cells_build(cells_t *cl, const box_t *b, const double *pos, int n, double rc) {
    cl->ok = 0;
    cl->start = NULL;
    cl->items = NULL;
    for (int k = 0; k < 3; ++k) {
        if (!(b->L[k] > 0.0) || !(rc > 0.0)) return;
        /* 1e-9 relative slack: an atom at distance rc is always in the 27-cell stencil */
        double e = b->L[k] / (rc * (1.0 + 1e-9));
        int c = (e > 1024.0) ? 1024 : (int)floor(e);
        if (c < 3) return;
        cl->nc[k] = c;
    }
    long ncell = (long)cl->nc[0] * cl->nc[1] * cl->nc[2];
    cl->start = (int *)calloc((size_t)ncell + 1, sizeof(int));
    cl->items = (int *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int));
    int *cid = (int *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int));
    int c3[3];
    for (int i = 0; i < n; ++i) {
        cid[i] = cell_of(b, cl->nc, pos + 3 * (size_t)i, c3);
        cl->start[cid[i] + 1]++;
    }
    for (long c = 0; c < ncell; ++c) cl->start[c + 1] += cl->start[c];
    int *fill = (int *)malloc((size_t)ncell * sizeof(int));
    memcpy(fill, cl->start, (size_t)ncell * sizeof(int));
    for (int i = 0; i < n; ++i) cl->items[fill[cid[i]]++] = i;
    free(fill);
    free(cid);
    cl->ok = 1;
}

static void cells_free(cells_t *cl) {
    free(cl->start);
    free(cl->items);
}

static int cmp_int(const void *a, const void *b) {
    int x = *(const int *)a, y = *(const int *)b;
    return (x > y) - (x < y);
}

/* Neighbour indices j (ascending) of centre r with lowsq < r2 <= highsq.  Returns the count; writes at
 * most cap indices. */
static int gather_neighbors(const cells_t *cl, const box_t *b, const double *pos, int n, const double *r,
                            double lowsq, double highsq, int *out, int cap) {
    int cnt = 0;
    double d[3];
    if (!cl->ok) {
        for (int j = 0; j < n; ++j) {
            min_image(b, pos + 3 * (size_t)j, r, d);
            double s = sumsq(d);
            if (s > lowsq && s <= highsq) {
                if (cnt < cap) out[cnt] = j;
                ++cnt;
            }
        }
        return cnt;
    }
    int c3[3];
    cell_of(b, cl->nc, r, c3);
    for (int dz = -1; dz <= 1; ++dz)
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                int cx = (c3[0] + dx + cl->nc[0]) % cl->nc[0];
                int cy = (c3[1] + dy + cl->nc[1]) % cl->nc[1];
                int cz = (c3[2] + dz + cl->nc[2]) % cl->nc[2];
                int c = (cz * cl->nc[1] + cy) * cl->nc[0] + cx;
                for (int t = cl->start[c]; t < cl->start[c + 1]; ++t) {
                    int j = cl->items[t];
                    min_image(b, pos + 3 * (size_t)j, r, d);
                    double s = sumsq(d);
                    if (s > lowsq && s <= highsq) {
                        if (cnt < cap) out[cnt] = j;
                        ++cnt;
                    }
                }
            }
    if (cnt <= cap) qsort(out, (size_t)cnt, sizeof(int), cmp_int);
    return cnt;
}

#define WOL_MAXNB 4096

/* ------------------------------------------------------------------------------------------ */

/* np.histogram(x, bins=nbins, range=[lo, hi]) uniform-bin rule (numpy/lib/_histograms_impl.py) */
static int hist_bin(double x, double lo, double hi, int nbins) {
    if (!(x >= lo) || !(x <= hi)) return -1;
    double denom = hi - lo;
    double f = ((x - lo) / denom) * (double)nbins;
    long idx = (long)f;
    if (idx == nbins) idx -= 1;
    double step = denom / (double)nbins;
    /* edges = linspace(lo, hi, nbins+1): k*step + lo, last forced to hi */
    double e_lo = (idx == nbins) ? hi : (double)idx * step + lo;
    if (x < e_lo) {
        idx -= 1;
    } else if (idx != nbins - 1) {
        double e_hi = (idx + 1 == nbins) ? hi : (double)(idx + 1) * step + lo;
        if (x >= e_hi) idx += 1;
    }
    return (int)idx;
}

int wol_oracle_histogram(const double *x, int64_t n, double lo, double hi, int nbins, int64_t *hist) {
    for (int64_t i = 0; i < n; ++i) {
        int b = hist_bin(x[i], lo, hi, nbins);
        if (b >= 0) hist[b]++;
    }
    return 0;
}

/* getCosAngs (water_properties.py:210-250) + the histogram / tetrahedral-window sums of
 * tetrahedralMetrics (:328-335).
 *   ncount[m]      neighbour count per centre (the reference's numAngs)
 *   ang_out        optional, all angles in reference order, capacity ang_cap
 *   n_ang_out      total number of angles
 *   hist[nbins]    accumulated (+=) angle histogram on [hlo, hhi]
 *   tet[3]         accumulated: count of 100<=ang<=120, sum cos(ang*pi/180), sum cos^2
 */
int wol_oracle_three_body(const double *sub, int m, const double *pos, int n, const double *boxl,
                          double lowcut, double highcut, int32_t *ncount, double *ang_out, int64_t ang_cap,
                          int64_t *n_ang_out, int64_t *hist, int nbins, double hlo, double hhi, double *tet) {
    box_t b;
    box_init(&b, boxl);
    cells_t cl;
    cells_build(&cl, &b, pos, n, highcut);
    double lowsq = lowcut * lowcut, highsq = highcut * highcut;
    int64_t na = 0;
    int rc = 0;
    /* Centres are independent; integer histogram sums commute, so threads keep private bins and
     * merge.  The materialised angle list has a defined order, so that variant runs on one thread. */
#pragma omp parallel if (ang_out == NULL && m > 4096)
    {
        int *nb = (int *)malloc(WOL_MAXNB * sizeof(int));
        double *img = (double *)malloc(WOL_MAXNB * 3 * sizeof(double));
        int64_t *lh = hist ? (int64_t *)calloc((size_t)nbins, sizeof(int64_t)) : NULL;
        double lt[3] = {0.0, 0.0, 0.0};
        int64_t lna = 0;
#pragma omp for schedule(static)
        for (int i = 0; i < m; ++i) {
            const double *r = sub + 3 * (size_t)i;
            int k = gather_neighbors(&cl, &b, pos, n, r, lowsq, highsq, nb, WOL_MAXNB);
            if (k > WOL_MAXNB) {
#pragma omp atomic write
                rc = 2;
                continue;
            }
            if (ncount) ncount[i] = k;
            /* tetraCosAng: distvec = neigh - ref, re-imaged, then ref + distvec (waterlib.f90:880-883) */
            for (int a = 0; a < k; ++a) {
                double d[3];
                min_image(&b, pos + 3 * (size_t)nb[a], r, d);
                for (int c = 0; c < 3; ++c) img[3 * a + c] = r[c] + d[c];
            }
            for (int a = 0; a < k; ++a)
                for (int c2 = a + 1; c2 < k; ++c2) {
                    double ang = cos_angle3(img + 3 * a, r, img + 3 * c2);
                    if (ang_out && lna < ang_cap) ang_out[lna] = ang;
                    ++lna;
                    if (lh) {
                        int bin = hist_bin(ang, hlo, hhi, nbins);
                        if (bin >= 0) lh[bin]++;
                    }
                    if (tet && ang >= 100.0 && ang <= 120.0) {
                        double c = cos(ang * kPi / 180.0);
                        lt[0] += 1.0;
                        lt[1] += c;
                        lt[2] += c * c;
                    }
                }
        }
#pragma omp critical
        {
            na += lna;
            if (lh) for (int k2 = 0; k2 < nbins; ++k2) hist[k2] += lh[k2];
            if (tet) for (int k2 = 0; k2 < 3; ++k2) tet[k2] += lt[k2];
        }
        free(lh);
        free(nb);
        free(img);
    }
    if (n_ang_out) *n_ang_out = na;
    cells_free(&cl);
    return rc;
}

typedef struct {
    double dist;
    int idx;
} cand_t;

/* stable ordering: distance, then candidate position (== ascending atom index) */
static int cmp_cand(const void *a, const void *b) {
    const cand_t *x = (const cand_t *)a, *y = (const cand_t *)b;
    if (x->dist < y->dist) return -1;
    if (x->dist > y->dist) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx);
}

/* getOrderParamq (water_properties.py:344-391).
 *   q[m], nn4[m*4] (atom indices of the selected neighbours in selection order, -1 padded),
 *   ncount[m] = number of neighbours within (lowcut, highcut].
 * Tie rule: equal distances -> smaller atom index first (np.argsort(kind='stable') on the
 * ascending-index candidate list). */
int wol_oracle_order_param_q(const double *sub, int m, const double *pos, int n, const double *boxl,
                             double lowcut, double highcut, double *q, int32_t *nn4, int32_t *ncount) {
    box_t b;
    box_init(&b, boxl);
    cells_t cl;
    cells_build(&cl, &b, pos, n, highcut);
    double lowsq = lowcut * lowcut, highsq = highcut * highcut;
    int cap = n > 0 ? (n < 65536 ? n : 65536) : 1;
    int rc = 0;
#pragma omp parallel if (m > 4096)
    {
        int *nb = (int *)malloc((size_t)cap * sizeof(int));
        cand_t *cd = (cand_t *)malloc((size_t)cap * sizeof(cand_t));
        double *img = (double *)malloc((size_t)cap * 3 * sizeof(double));
#pragma omp for schedule(static)
        for (int i = 0; i < m; ++i) {
            const double *r = sub + 3 * (size_t)i;
            int k = gather_neighbors(&cl, &b, pos, n, r, lowsq, highsq, nb, cap);
            if (k > cap) {
#pragma omp atomic write
                rc = 2;
                continue;
            }
            if (ncount) ncount[i] = k;
            if (nn4) for (int a = 0; a < 4; ++a) nn4[4 * (size_t)i + a] = -1;
            q[i] = 0.0;
            if (k == 0) continue;
            /* reimage (waterlib.f90:43-45) then np.linalg.norm(thisPos - apos) */
            for (int a = 0; a < k; ++a) {
                double d[3], e[3];
                min_image(&b, pos + 3 * (size_t)nb[a], r, d);
                for (int c = 0; c < 3; ++c) {
                    img[3 * a + c] = r[c] + d[c];
                    e[c] = img[3 * a + c] - r[c];
                }
                cd[a].dist = sqrt(sumsq(e));
                cd[a].idx = a;
            }
            qsort(cd, (size_t)k, sizeof(cand_t), cmp_cand);
            int k4 = k < 4 ? k : 4;
            double p4[4][3];
            for (int a = 0; a < k4; ++a) {
                if (nn4) nn4[4 * (size_t)i + a] = nb[cd[a].idx];
                /* tetraCosAng re-images the already re-imaged position again (waterlib.f90:880-883) */
                double d[3];
                min_image(&b, img + 3 * cd[a].idx, r, d);
                for (int c = 0; c < 3; ++c) p4[a][c] = r[c] + d[c];
            }
            double ang[6];
            int na = 0;
            for (int a = 0; a < k4; ++a)
                for (int c2 = a + 1; c2 < k4; ++c2) ang[na++] = cos_angle3(p4[a], r, p4[c2]);
            /* padding with 180 deg when fewer than 4 neighbours (water_properties.py:379-384) */
            if (k == 1) {
                na = 0;
                for (int a = 0; a < 6; ++a) ang[na++] = 180.0;
            } else if (k == 2) {
                for (int a = 0; a < 5; ++a) ang[na++] = 180.0;
            } else if (k == 3) {
                for (int a = 0; a < 3; ++a) ang[na++] = 180.0;
            }
            double s = 0.0;
            for (int a = 0; a < na; ++a) {
                double t = cos(ang[a] * kPi / 180.0) + (1.0 / 3.0);
                s += t * t;
            }
            q[i] = 1.0 - (3.0 / 8.0) * s;
        }
        free(nb);
        free(cd);
        free(img);
    }
    cells_free(&cl);
    return rc;
}

/* Dense neighbour matrix (m x n, row-major int32), nearNeighbors / allNearNeighbors semantics. */
int wol_oracle_neighbor_matrix(const double *sub, int m, const double *pos, int n, const double *boxl,
                               double lowcut, double highcut, int32_t *mat) {
    box_t b;
    box_init(&b, boxl);
    double lowsq = lowcut * lowcut, highsq = highcut * highcut;
    double d[3];
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) {
            min_image(&b, pos + 3 * (size_t)j, sub + 3 * (size_t)i, d);
            double s = sumsq(d);
            mat[(size_t)i * n + j] = (s > lowsq && s <= highsq) ? 1 : 0;
        }
    return 0;
}

/* Shell selection (orderParam_lib.py:495-498): mask[j] = 1 iff some solute atom i has
 * lowcut^2 < r2(i,j) <= cutoff^2. */
int wol_oracle_shell_mask(const double *sol, int ns, const double *wat, int nw, const double *boxl,
                          double lowcut, double cutoff, int32_t *mask) {
    box_t b;
    box_init(&b, boxl);
    double lowsq = lowcut * lowcut, highsq = cutoff * cutoff;
    double d[3];
    for (int j = 0; j < nw; ++j) {
        mask[j] = 0;
        for (int i = 0; i < ns; ++i) {
            min_image(&b, wat + 3 * (size_t)j, sol + 3 * (size_t)i, d);
            double s = sumsq(d);
            if (s > lowsq && s <= highsq) { mask[j] = 1; break; }
        }
    }
    return 0;
}

/* generalHbonds (waterlib.f90:1156-1210) reduced to the sums hbCalc takes
 * (orderParam_lib.py:867-884): acc_count[i] = sum_j bond(i,j), don_count[j] = sum_i bond(i,j);
 * optional dense matrix (na x nd row-major). */
int wol_oracle_hbonds(const double *acc, int na, const double *don, const double *donh, int nd,
                      const double *boxl, double distcut, double angcut, int32_t *acc_count, int32_t *don_count,
                      int32_t *mat) {
    box_t b;
    box_init(&b, boxl);
    double cutsq = distcut * distcut;
    const double tiny = (double)1.0e-2f; /* the Fortran literal 1.0E-2 is single precision (:1187) */
    cells_t cl;
    cells_build(&cl, &b, don, nd, distcut);
    if (acc_count) memset(acc_count, 0, (size_t)na * sizeof(int32_t));
    if (don_count) memset(don_count, 0, (size_t)nd * sizeof(int32_t));
    if (mat) memset(mat, 0, (size_t)na * nd * sizeof(int32_t));
    int cap = nd > 0 ? nd : 1;
    int *nb = (int *)malloc((size_t)cap * sizeof(int));
    for (int i = 0; i < na; ++i) {
        const double *pa = acc + 3 * (size_t)i;
        /* candidates: distSq <= distCutSq and distSq > 1.0E-2 */
        int k = gather_neighbors(&cl, &b, don, nd, pa, tiny, cutsq, nb, cap);
        for (int t = 0; t < k; ++t) {
            int j = nb[t];
            const double *ph = donh + 3 * (size_t)j;
            double av[3], dv[3];
            min_image(&b, pa, ph, av);
            double an = sqrt(sumsq(av));
            for (int c = 0; c < 3; ++c) av[c] = av[c] / an;
            min_image(&b, don + 3 * (size_t)j, ph, dv);
            double dn = sqrt(sumsq(dv));
            for (int c = 0; c < 3; ++c) dv[c] = dv[c] / dn;
            double ang = ang_between(av, dv);
            if (ang < angcut) continue;
            if (acc_count) acc_count[i]++;
            if (don_count) don_count[j]++;
            if (mat) mat[(size_t)i * nd + j] = 1;
        }
    }
    free(nb);
    cells_free(&cl);
    return 0;
}

/* The tail of CosAngle3 (waterlib.f90:698-702) for an array of already-formed cosine ratios: clamp,
 * acos, the mod/branch, degrees.  Used to pin the device's threshold-table binning. */
int wol_oracle_angles_from_cos(const double *c, int64_t n, double *ang) {
    for (int64_t i = 0; i < n; ++i) {
        double phi = acos(fmin(1.0, fmax(-1.0, c[i])));
        double a = fmod(phi + kPi, kTwoPi) - kPi;
        if (a < -kPi) a += kTwoPi;
        ang[i] = a * kDegPerRad;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Slab / interface routines and the all-Fortran triplet histogram.                            */

/* WillardDensityField / WillardDensityPoints (waterlib.f90:1286-1341, :1351-1398): truncated, shifted
 * Gaussian density and its normalised gradient at `npts` points.  Sum over waters in index order, exactly
 * as the reference loops (every water is tested; the cut-off only zeroes its term). */
int wol_oracle_willard_points(const double *pos, int n, const double *pts, int64_t npts, const double *boxl,
                              double smoothlen, double *densvals, double *densnorms) {
    box_t b;
    box_init(&b, boxl);
    const double s2 = smoothlen * smoothlen;
    const double pref = pow(2.0 * kPi * s2, 1.5);
    const double shiftterm = exp(-9.0 / 2.0) / pref;
    const double cut = 9.0 * (smoothlen * smoothlen);
#pragma omp parallel for schedule(static) if (npts > 256)
    for (int64_t g = 0; g < npts; ++g) {
        const double *a = pts + 3 * (size_t)g;
        double dens = 0.0, nv[3] = {0.0, 0.0, 0.0};
        for (int l = 0; l < n; ++l) {
            double v[3];
            min_image(&b, a, pos + 3 * (size_t)l, v); /* thisvec = apos - watpos */
            double r2 = sumsq(v);
            if (r2 >= cut) continue; /* adds 0.0: no effect on the sums */
            double expterm = -r2 / (2.0 * s2);
            double densfunc = exp(expterm) / pref - shiftterm;
            for (int c = 0; c < 3; ++c) nv[c] = nv[c] + (-v[c] * (densfunc + shiftterm) / s2);
            dens = dens + densfunc;
        }
        densvals[g] = dens;
        double nn = sqrt(sumsq(nv));
        for (int c = 0; c < 3; ++c) densnorms[3 * (size_t)g + c] = nv[c] / nn;
    }
    return 0;
}

/* InterfaceWater (waterlib.f90:1414-1469).  Indices are 0-based here, -1 = no interface point (water) inside
 * distance^2 < 1000; such waters get allwatdists = 0 (the Fortran reads an uninitialised index there). */
int wol_oracle_interface_water(const double *pos, int n, const double *gridpos, const double *gridnorm, int ng,
                               double cutoff, const double *boxl, int32_t *watclose, int32_t *surfclose,
                               int32_t *numwater, double *allwatdists) {
    box_t b;
    box_init(&b, boxl);
    double *griddists = (double *)malloc((size_t)(ng > 0 ? ng : 1) * sizeof(double));
    for (int j = 0; j < ng; ++j) {
        griddists[j] = 1000.0;
        surfclose[j] = -1;
    }
    int count = 0;
    for (int i = 0; i < n; ++i) {
        const double *w = pos + 3 * (size_t)i;
        double watdist = 1000.0, d[3];
        int close = -1;
        for (int j = 0; j < ng; ++j) {
            min_image(&b, w, gridpos + 3 * (size_t)j, d);
            double s = sumsq(d);
            if (s < watdist) {
                close = j;
                watdist = s;
            }
            if (s < griddists[j]) {
                surfclose[j] = i;
                griddists[j] = s;
            }
        }
        watclose[i] = close;
        double proj = 0.0;
        if (close >= 0) {
            const double *cn = gridnorm + 3 * (size_t)close;
            min_image(&b, w, gridpos + 3 * (size_t)close, d);
            proj = (d[0] * cn[0] + d[1] * cn[1]) + d[2] * cn[2];
            if (proj <= cutoff) ++count;
        }
        allwatdists[i] = proj;
    }
    *numwater = count;
    free(griddists);
    return 0;
}

/* histrr3b (waterlib.f90:1550-1593): triplet histogram with ceiling binning.  hist[d1][d2][a] (0-based, C
 * order) counts; triplets whose bin index would be 0 or negative in the Fortran (coincident atoms, the 0 and
 * -180 degree returns of CosAngle3) write out of bounds there and are skipped here. */
int wol_oracle_histrr3b(const double *pos, int n, const double *boxl, double dwidth, int dnum, double awidth, int anum,
                        int64_t *hist) {
    box_t b;
    box_init(&b, boxl);
    const double zero[3] = {0.0, 0.0, 0.0};
    double *vec = (double *)malloc((size_t)(n > 0 ? n : 1) * 3 * sizeof(double));
    int *bin = (int *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int));
    for (int i = 0; i < n; ++i) {
        const double *r = pos + 3 * (size_t)i;
        for (int j = 0; j < n; ++j) {
            min_image(&b, pos + 3 * (size_t)j, r, vec + 3 * (size_t)j);
            bin[j] = (int)ceil(sqrt(sumsq(vec + 3 * (size_t)j)) / dwidth);
        }
        for (int j = 0; j < n; ++j) {
            if (j == i || bin[j] > dnum) continue;
            for (int k = j + 1; k < n; ++k) {
                if (k == i || bin[k] > dnum) continue;
                double ang = cos_angle3(vec + 3 * (size_t)j, zero, vec + 3 * (size_t)k);
                int ab = (int)ceil(ang / awidth);
                if (ab > anum) continue;
                if (bin[j] < 1 || bin[k] < 1 || ab < 1) continue;
                hist[((size_t)(bin[j] - 1) * dnum + (bin[k] - 1)) * anum + (ab - 1)]++;
            }
        }
    }
    free(vec);
    free(bin);
    return 0;
}

/* getLSI (structureLibs/water_properties.py:252-311): local structure index per centre.
 *   near  = atoms with lowcut^2 < r2 <= highcut^2 (minimum image), next = highcut^2 < r2 <= (highcut+3.7)^2
 *   a centre gets a value only if it has > 1 near and >= 1 next neighbours; the next neighbour taken is the one
 *   with the smallest NON-periodic distance sqrt(sum((Pos - apos)**2)) (:289, first index on ties); the sorted
 *   minimum-image distances (lsiDists, waterlib.f90:900-918) of near + that one give deltas whose population
 *   variance is the LSI.  lsi[i] is written only where has[i] = 1; num[i] = number of deltas (= near count). */
static int cmp_double(const void *a, const void *b) {
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

int wol_oracle_lsi(const double *sub, int m, const double *pos, int n, const double *boxl, double lowcut, double highcut,
                   double *lsi, int32_t *num, int32_t *has) {
    box_t b;
    box_init(&b, boxl);
    const double lowsq = lowcut * lowcut, highsq = highcut * highcut;
    const double nextsq = (highcut + 3.7) * (highcut + 3.7);
    double *dist = (double *)malloc((size_t)(n + 1) * sizeof(double));
    for (int i = 0; i < m; ++i) {
        const double *r = sub + 3 * (size_t)i;
        int k = 0, next = -1, n_next = 0;
        double next_raw = 0.0, d[3];
        for (int j = 0; j < n; ++j) {
            const double *p = pos + 3 * (size_t)j;
            min_image(&b, p, r, d);
            double s = sumsq(d);
            if (s > lowsq && s <= highsq) {
                dist[k++] = sqrt(s);
            } else if (s > highsq && s <= nextsq) {
                double e0 = p[0] - r[0], e1 = p[1] - r[1], e2 = p[2] - r[2];
                double raw = sqrt((e0 * e0 + e1 * e1) + e2 * e2);
                if (n_next == 0 || raw < next_raw) {
                    next_raw = raw;
                    next = j;
                }
                ++n_next;
            }
        }
        num[i] = 0;
        has[i] = 0;
        lsi[i] = 0.0;
        if (k > 1 && n_next > 0) {
            min_image(&b, pos + 3 * (size_t)next, r, d);
            dist[k++] = sqrt(sumsq(d));
            qsort(dist, (size_t)k, sizeof(double), cmp_double);
            int nd = k - 1;
            double mean = 0.0;
            for (int t = 0; t < nd; ++t) mean += dist[t + 1] - dist[t];
            mean /= (double)nd;
            double var = 0.0;
            for (int t = 0; t < nd; ++t) {
                double x = (dist[t + 1] - dist[t]) - mean;
                var += x * x;
            }
            lsi[i] = var / (double)nd;
            num[i] = nd;
            has[i] = 1;
        }
    }
    free(dist);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Pair-distance histograms: RadialDist / RadialDistSame / PairDistanceHistogram (waterlib.f90:193-231, :316-353,
 * :358-389).  counts[k], k = 0-based bin of nbin = ceiling(dist / binwidth); distances that index bin 0 in the
 * Fortran (dist == 0, an out-of-bounds write there) are skipped.
 *   mode 0  RadialDist:            every (i in pos2, j in pos1) pair
 *   mode 1  RadialDistSame:        pairs i < j of pos1 (pos2 ignored)
 *   mode 2  PairDistanceHistogram: every (i in pos1, j in pos2) pair, dist == 0 skipped */
int wol_oracle_pair_hist(int mode, const double *pos1, int n1, const double *pos2, int n2, const double *boxl,
                         double binwidth, int totbins, int64_t *counts) {
    box_t b;
    box_init(&b, boxl);
    const double *outer = (mode == 0) ? pos2 : pos1, *inner = (mode == 2) ? pos2 : pos1;
    const int no = (mode == 0) ? n2 : n1, ni = (mode == 2) ? n2 : n1;
    double d[3];
    for (int i = 0; i < no; ++i)
        for (int j = (mode == 1) ? i + 1 : 0; j < ni; ++j) {
            min_image(&b, inner + 3 * (size_t)j, outer + 3 * (size_t)i, d); /* distVec = jPos - iPos */
            double dist = sqrt(sumsq(d));
            double nb = ceil(dist / binwidth);
            if (nb >= 1.0 && nb <= (double)totbins) counts[(int)nb - 1]++;
        }
    return 0;
}

/* getOrderParamPsi (structureLibs/water_properties.py:393-433).  Intended: psi_i = | mean over neighbour pairs of
 * exp(6 i theta) |.  ACTUAL: the complex mean is stored into a float64 array (:428), which discards its imaginary
 * part (numpy ComplexWarning), so what the reference returns is | mean cos(6 theta) |.  That is what is restated.
 * theta from tetraCosAng/CosAngle3 (degrees, then * pi / 180), neighbours inside (lowcut, highcut]; 0 when the
 * centre has fewer than two neighbours.  The reference sorts the neighbours by distance first; the mean does not
 * depend on the order beyond rounding, ascending index is used here. */
int wol_oracle_psi(const double *sub, int m, const double *pos, int n, const double *boxl, double lowcut, double highcut,
                   double *psi) {
    box_t b;
    box_init(&b, boxl);
    const double lowsq = lowcut * lowcut, highsq = highcut * highcut;
    double *img = (double *)malloc((size_t)(n > 0 ? n : 1) * 3 * sizeof(double));
    for (int i = 0; i < m; ++i) {
        const double *r = sub + 3 * (size_t)i;
        int k = 0;
        double d[3];
        for (int j = 0; j < n; ++j) {
            min_image(&b, pos + 3 * (size_t)j, r, d);
            double s = sumsq(d);
            if (s > lowsq && s <= highsq) {
                /* reimage (ref + d), then tetraCosAng reimages again: ref + minimg((ref + d) - ref) */
                double p[3], d2[3];
                for (int c = 0; c < 3; ++c) p[c] = r[c] + d[c];
                min_image(&b, p, r, d2);
                for (int c = 0; c < 3; ++c) img[3 * k + c] = r[c] + d2[c];
                ++k;
            }
        }
        psi[i] = 0.0;
        if (k > 1) {
            double re = 0.0;
            long np_ = 0;
            for (int a = 0; a < k; ++a)
                for (int c2 = a + 1; c2 < k; ++c2) {
                    double ang = cos_angle3(img + 3 * a, r, img + 3 * c2) * kPi / 180.0;
                    re += cos(6.0 * ang);
                    ++np_;
                }
            re /= (double)np_;
            psi[i] = sqrt(re * re);
        }
    }
    free(img);
    return 0;
}

/* DensityField (waterlib.f90:1219-1268): number of waters inside the cube of edge binwidth = gridx(2) - gridx(1)
 * centred on each grid point (inclusive faces, minimum image), divided by binwidth**3.0.  densvals[nx][ny][nz]. */
int wol_oracle_density_field(const double *pos, int n, const double *gx, int nx, const double *gy, int ny, const double *gz,
                             int nz, const double *boxl, double *densvals) {
    box_t b;
    box_init(&b, boxl);
    const double binwidth = gx[1] - gx[0];
    const double h = binwidth / 2.0, vol = pow(binwidth, 3.0);
    for (int i = 0; i < nx; ++i)
        for (int j = 0; j < ny; ++j)
            for (int k = 0; k < nz; ++k) {
                const double a[3] = {gx[i], gy[j], gz[k]};
                double dens = 0.0, v[3];
                for (int l = 0; l < n; ++l) {
                    min_image(&b, pos + 3 * (size_t)l, a, v); /* thisvec = watpos - apos */
                    int in = 1;
                    for (int c = 0; c < 3; ++c) {
                        const double w = a[c] + v[c];
                        if (w < (a[c] - h) || w > (a[c] + h)) in = 0;
                    }
                    if (in) dens = dens + 1.0;
                }
                densvals[((size_t)i * ny + j) * nz + k] = dens / vol;
            }
    return 0;
}

/* watOrient (fortran/waterlib.f90:973-1011): per water the angles (degrees, AngBetween) between refvec and the dipole
 * direction (sum of the two minimum-imaged O->H vectors, imaged once more) and between refvec and the normal of the
 * molecular plane (cross product of the O->H vectors). hpos holds H1, H2 of water i at rows 2i, 2i+1. */
int wol_oracle_watorient(const double *opos, int no, const double *hpos, const double *refvec, const double *boxl, double *angdip,
                         double *angplane) {
    box_t b;
    box_init(&b, boxl);
    const double rn = sqrt((refvec[0] * refvec[0] + refvec[1] * refvec[1]) + refvec[2] * refvec[2]);
    const double ref[3] = {refvec[0] / rn, refvec[1] / rn, refvec[2] / rn};
    for (int i = 0; i < no; ++i) {
        double v1[3], v2[3], dip[3], pl[3], u[3];
        min_image(&b, hpos + 3 * (size_t)(2 * i), opos + 3 * (size_t)i, v1);
        min_image(&b, hpos + 3 * (size_t)(2 * i + 1), opos + 3 * (size_t)i, v2);
        for (int k = 0; k < 3; ++k) {
            const double t = v1[k] + v2[k];
            dip[k] = t - b.L[k] * round(t * b.iL[k]);
        }
        double n = sqrt(sumsq(dip));
        for (int k = 0; k < 3; ++k) u[k] = dip[k] / n;
        angdip[i] = ang_between(u, ref);
        pl[0] = v1[1] * v2[2] - v1[2] * v2[1]; /* crossProd3, waterlib.f90:26-28 */
        pl[1] = v1[2] * v2[0] - v1[0] * v2[2];
        pl[2] = v1[0] * v2[1] - v1[1] * v2[0];
        n = sqrt(sumsq(pl));
        for (int k = 0; k < 3; ++k) u[k] = pl[k] / n;
        angplane[i] = ang_between(u, ref);
    }
    return 0;
}

/* binOnGrid (fortran/waterlib.f90:1047-1099): atoms per cubic bin (left edge inclusive), counted only inside the sphere
 * of diameter binwidth centred in the bin; no periodic wrap. outhist [nx-1][ny-1][nz-1] row-major. */
int wol_oracle_binongrid(const double *opos, int n, const double *xb, int nx, const double *yb, int ny, const double *zb, int nz,
                         int32_t *outhist) {
    const double binwidth = xb[1] - xb[0];
    if (yb[1] - yb[0] != binwidth || zb[1] - zb[0] != binwidth) return -1;
    const double radsq = binwidth * binwidth / 4.0;
    for (size_t i = 0; i < (size_t)(nx - 1) * (ny - 1) * (nz - 1); ++i) outhist[i] = 0;
    for (int i = 0; i < n; ++i) {
        const double *p = opos + 3 * (size_t)i;
        const double fx = floor((p[0] - xb[0]) / binwidth), fy = floor((p[1] - yb[0]) / binwidth), fz = floor((p[2] - zb[0]) / binwidth);
        if (!(fx >= 0 && fx < nx - 1) || !(fy >= 0 && fy < ny - 1) || !(fz >= 0 && fz < nz - 1)) continue;
        const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
        const double v[3] = {p[0] - (xb[ix] + binwidth * 0.5), p[1] - (yb[iy] + binwidth * 0.5), p[2] - (zb[iz] + binwidth * 0.5)};
        if (sumsq(v) <= radsq) outhist[((size_t)ix * (ny - 1) + iy) * (nz - 1) + iz] += 1;
    }
    return 0;
}
