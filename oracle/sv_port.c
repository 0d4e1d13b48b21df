/*
 * sv_port.c -- plain C restatement of the reference's CPU algorithm for the hot path.
 * TEST / BASELINE INFRASTRUCTURE ONLY: used by tests/ (against the numpy oracle) and by
 * bench.py's cpu_baseline / --impl reference legs.  The product never links it.
 *
 * What it restates: the reference applies a circuit gate by gate -- one
 * tn.contract_between -> backend.tensordot(gate, state) per contraction-path step once the
 * running tensor has 2^n entries (tensorcircuit/cons.py:605-623), i.e. one full read + write of
 * the state per recorded gate, no fusion of multi-qubit gates
 * (tensorcircuit/cons.py:236-279 only pre-merges single-qubit gates into a neighbour).
 * svp_apply() is that step: psi'[.. o ..] = sum_i U[o, i] psi[.. i ..] over the 2^k-amplitude
 * groups (axis convention of tensorcircuit/basecircuit.py:213-215; bit j of the matrix index is
 * amplitude-index bit bits[j], ascending), parallelised over groups with OpenMP so that it can
 * use every host core like the jax/BLAS CPU paths do.  Pauli expectation and the CDF sampler
 * restate tensorcircuit/quantum.py:1461-1482 and
 * tensorcircuit/backends/abstract_backend.py:1145-1157.
 *
 * Pinned by tests/test_oracle_port.py against oracle/tc_oracle.py (itself pinned against the
 * reference's golden test values).
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXK 5

static inline uint64_t insert_zeros(uint64_t g, const int* bits, int k) {
    for (int j = 0; j < k; ++j) {
        const uint64_t lo = g & ((1ull << bits[j]) - 1ull);
        g = ((g >> bits[j]) << (bits[j] + 1)) | lo;
    }
    return g;
}

/* torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU baseline is meant to use the host */
void svp_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int svp_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

#define DEFINE_APPLY(NAME, CT)                                                                 \
    int NAME(CT* psi, int n, int k, const int* bits, const double complex* u) {                \
        if (k < 1 || k > MAXK || k > n) return -1;                                             \
        const int d = 1 << k;                                                                  \
        uint64_t off[1 << MAXK];                                                               \
        CT m[(1 << MAXK) * (1 << MAXK)];                                                       \
        for (int j = 0; j < d; ++j) {                                                          \
            off[j] = 0;                                                                        \
            for (int b = 0; b < k; ++b)                                                        \
                if ((j >> b) & 1) off[j] |= 1ull << bits[b];                                   \
        }                                                                                      \
        for (int i = 0; i < d * d; ++i) m[i] = (CT)u[i];                                       \
        const int64_t ng = (int64_t)1 << (n - k);                                              \
        _Pragma("omp parallel for schedule(static)") for (int64_t g = 0; g < ng; ++g) {        \
            const uint64_t base = insert_zeros((uint64_t)g, bits, k);                          \
            CT v[1 << MAXK];                                                                   \
            for (int j = 0; j < d; ++j) v[j] = psi[base | off[j]];                             \
            for (int i = 0; i < d; ++i) {                                                      \
                CT acc = 0;                                                                    \
                for (int j = 0; j < d; ++j) acc += m[i * d + j] * v[j];                        \
                psi[base | off[i]] = acc;                                                      \
            }                                                                                  \
        }                                                                                      \
        return 0;                                                                              \
    }

DEFINE_APPLY(svp_apply_c64, float complex)
DEFINE_APPLY(svp_apply_c128, double complex)

void svp_init_zero_c64(float complex* psi, int n) {
    const int64_t N = (int64_t)1 << n;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) psi[i] = 0;
    psi[0] = 1;
}

void svp_init_zero_c128(double complex* psi, int n) {
    const int64_t N = (int64_t)1 << n;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) psi[i] = 0;
    psi[0] = 1;
}

/* <psi| P |psi> = sum_r conj(psi_r) (-1)^{popc(r & sign)} (-i)^{ny} psi_{r ^ flip} */
#define DEFINE_EXPECT(NAME, CT)                                                                \
    void NAME(const CT* psi, int n, uint64_t flip, uint64_t sign, int ny, double* out) {       \
        const int64_t N = (int64_t)1 << n;                                                     \
        double re = 0, im = 0;                                                                 \
        _Pragma("omp parallel for schedule(static) reduction(+ : re, im)")                     \
        for (int64_t r = 0; r < N; ++r) {                                                      \
            const double complex t = conj((double complex)psi[r]) * (double complex)psi[r ^ flip]; \
            const int neg = __builtin_popcountll((uint64_t)r & sign) & 1;                      \
            re += neg ? -creal(t) : creal(t);                                                  \
            im += neg ? -cimag(t) : cimag(t);                                                  \
        }                                                                                      \
        double complex v = re + im * I;                                                        \
        for (int j = 0; j < (ny & 3); ++j) v *= -I;                                            \
        out[0] = creal(v);                                                                     \
        out[1] = cimag(v);                                                                     \
    }

DEFINE_EXPECT(svp_expect_c64, float complex)
DEFINE_EXPECT(svp_expect_c128, double complex)

/* p = |psi|^2; cdf = cumsum(p / sum p) in float64; idx = first i with cdf[i] >= cdf[-1]*(1-u) */
#define DEFINE_SAMPLE(NAME, CT)                                                                \
    int NAME(const CT* psi, int n, const double* u, int64_t shots, int64_t* out) {             \
        const int64_t N = (int64_t)1 << n;                                                     \
        double* cdf = (double*)malloc(sizeof(double) * (size_t)N);                             \
        if (!cdf) return -1;                                                                   \
        double tot = 0;                                                                        \
        for (int64_t i = 0; i < N; ++i) {                                                      \
            const double complex a = psi[i];                                                   \
            tot += creal(a) * creal(a) + cimag(a) * cimag(a);                                  \
        }                                                                                      \
        double run = 0;                                                                        \
        for (int64_t i = 0; i < N; ++i) {                                                      \
            const double complex a = psi[i];                                                   \
            run += (creal(a) * creal(a) + cimag(a) * cimag(a)) / tot;                          \
            cdf[i] = run;                                                                      \
        }                                                                                      \
        _Pragma("omp parallel for schedule(static)") for (int64_t s = 0; s < shots; ++s) {     \
            const double r = cdf[N - 1] * (1.0 - u[s]);                                        \
            int64_t lo = 0, hi = N;                                                            \
            while (lo < hi) {                                                                  \
                const int64_t mid = (lo + hi) >> 1;                                            \
                if (cdf[mid] < r) lo = mid + 1; else hi = mid;                                 \
            }                                                                                  \
            out[s] = lo;                                                                       \
        }                                                                                      \
        free(cdf);                                                                             \
        return 0;                                                                              \
    }

DEFINE_SAMPLE(svp_sample_c64, float complex)
DEFINE_SAMPLE(svp_sample_c128, double complex)
