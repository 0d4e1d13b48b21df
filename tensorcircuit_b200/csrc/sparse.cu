// Sparse-operator expectation  <psi| H |psi>  for an operator kept on the device in COO form.
//
// Replaces tensorcircuit/templates/measurements.py:173-188 (sparse_expectation: backend
// sparse_dense_matmul of the COO Hamiltonian with the ket, then <psi|tmp>) and the dense branch of
// operator_expectation (measurements.py:156-170) for operators that are NOT Pauli sums; Pauli sums
// built by quantum.PauliStringSum2COO never become a matrix here -- they stay strings and go
// through the Pauli kernels (expect.cu / xexpect.cu).
//
// One thread per stored element k:  acc += conj(psi[row_k]) * val_k * psi[col_k], float64
// accumulation, grid-stride loop over a fixed grid, fixed-order block and grid reductions (the
// result does not depend on timing).  The kernel is bound by the 32 bytes of (row, col, value) per
// element plus two gathered amplitudes; the operator is uploaded once and stays resident over the
// iterations of a VQE loop.
//
// The same file holds the two kernels of the adjoint-state gradient (autodiff.py; replaces the
// backend autodiff of tensorcircuit/backends/jax_backend.py:668-776 for losses built from
// expectation values):
//   * pauli_sum_kernel:  dst = sum_t w_t P_t src        (lambda = H psi, the seed of the backward sweep;
//                        closed form of the rows of quantum.py:1461-1482)
//   * transition_kernel: <bra| G_j |ket> for a list of local (<= 2-bit) operators G_j = dM_j/dtheta M_j^+,
//                        every operator of a launch evaluated on the same pair of states.
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace tcb {

__device__ __forceinline__ double sp_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

constexpr int COO_THREADS = 256;
constexpr int COO_MAX_CTAS = 148 * 8;

template <typename C>
__global__ void __launch_bounds__(COO_THREADS) coo_expect_kernel(const C* __restrict__ state, int nbits, long long nnz,
                                                                 const long long* __restrict__ rows, const long long* __restrict__ cols,
                                                                 const double2* __restrict__ vals, double* __restrict__ partials) {
    __shared__ double red[COO_THREADS / 32][2];
    const C* psi = state + ((uint64_t)blockIdx.y << nbits);
    double re = 0.0, im = 0.0;
    for (long long k = (long long)blockIdx.x * COO_THREADS + threadIdx.x; k < nnz; k += (long long)gridDim.x * COO_THREADS) {
        const C a = psi[rows[k]];
        const C b = psi[cols[k]];
        const double2 v = vals[k];
        const double tr = v.x * (double)b.x - v.y * (double)b.y;  // val * psi[col]
        const double ti = v.x * (double)b.y + v.y * (double)b.x;
        re += (double)a.x * tr + (double)a.y * ti;  // conj(psi[row]) * t
        im += (double)a.x * ti - (double)a.y * tr;
    }
    re = sp_warp_sum(re);
    im = sp_warp_sum(im);
    const int tid = threadIdx.x;
    if ((tid & 31) == 0) {
        red[tid >> 5][0] = re;
        red[tid >> 5][1] = im;
    }
    __syncthreads();
    if (tid == 0) {
        re = im = 0.0;
        for (int w = 0; w < COO_THREADS / 32; ++w) {
            re += red[w][0];
            im += red[w][1];
        }
        double* o = partials + ((uint64_t)blockIdx.y * gridDim.x + blockIdx.x) * 2;
        o[0] = re;
        o[1] = im;
    }
}

__global__ void __launch_bounds__(COO_THREADS) coo_final_kernel(const double* __restrict__ partials, int nctas, double* __restrict__ out) {
    __shared__ double red[COO_THREADS / 32][2];
    const double* src = partials + (uint64_t)blockIdx.x * nctas * 2;
    const int tid = threadIdx.x;
    double re = 0.0, im = 0.0;
    for (int c = tid; c < nctas; c += COO_THREADS) {
        re += src[2 * c];
        im += src[2 * c + 1];
    }
    re = sp_warp_sum(re);
    im = sp_warp_sum(im);
    if ((tid & 31) == 0) {
        red[tid >> 5][0] = re;
        red[tid >> 5][1] = im;
    }
    __syncthreads();
    if (tid == 0) {
        re = im = 0.0;
        for (int w = 0; w < COO_THREADS / 32; ++w) {
            re += red[w][0];
            im += red[w][1];
        }
        out[2 * blockIdx.x] = re;
        out[2 * blockIdx.x + 1] = im;
    }
}

static unsigned coo_grid(int64_t nnz) {
    int64_t g = (nnz + COO_THREADS - 1) / COO_THREADS;
    if (g < 1) g = 1;
    if (g > COO_MAX_CTAS) g = COO_MAX_CTAS;
    return (unsigned)g;
}


// ---- dst = sum_t coef_t (-1)^{parity(r & sign_t)} src[r ^ flip_t] -------------------------------------
// Terms arrive sorted by flip mask (host), so that strings with the same flip share one partner load;
// they are staged through shared memory in chunks.
constexpr int PS_CHUNK = 256;

struct PauliTerm {
    uint64_t flip, sign;
    double cr, ci;
};

template <typename C>
__global__ void __launch_bounds__(256) pauli_sum_kernel(const C* __restrict__ src, C* __restrict__ dst, int nbits, int nterms,
                                                        const PauliTerm* __restrict__ terms) {
    __shared__ PauliTerm sh[PS_CHUNK];
    const uint64_t r = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    const bool live = r < (1ull << nbits);
    const C* s = src + ((uint64_t)blockIdx.y << nbits);
    double ar = 0.0, ai = 0.0;
    for (int t0 = 0; t0 < nterms; t0 += PS_CHUNK) {
        const int nc = nterms - t0 < PS_CHUNK ? nterms - t0 : PS_CHUNK;
        __syncthreads();
        if ((int)threadIdx.x < nc) sh[threadIdx.x] = terms[t0 + threadIdx.x];
        __syncthreads();
        if (!live) continue;
        int t = 0;
        while (t < nc) {
            const uint64_t f = sh[t].flip;
            const C b = s[r ^ f];
            double cr = 0.0, ci = 0.0;
            do {
                const double sg = (__popcll(r & sh[t].sign) & 1) ? -1.0 : 1.0;
                cr += sg * sh[t].cr;
                ci += sg * sh[t].ci;
                ++t;
            } while (t < nc && sh[t].flip == f);
            ar += cr * (double)b.x - ci * (double)b.y;
            ai += cr * (double)b.y + ci * (double)b.x;
        }
    }
    if (live) {
        C o;
        o.x = (decltype(o.x))ar;
        o.y = (decltype(o.y))ai;
        dst[((uint64_t)blockIdx.y << nbits) + r] = o;
    }
}

// ---- dst (+)= coef * H src for a CSR operator: one thread per row, entries of a row summed in storage order ----
// (the seed lambda = H psi of the adjoint sweep when the Hamiltonian is a generic sparse matrix)
template <typename C>
__global__ void __launch_bounds__(256) csr_matvec_kernel(const C* __restrict__ src, C* __restrict__ dst, int nbits,
                                                         const long long* __restrict__ indptr, const long long* __restrict__ indices,
                                                         const double2* __restrict__ vals, double cr, double ci, int accumulate) {
    const uint64_t r = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (r >= (1ull << nbits)) return;
    double ar = 0.0, ai = 0.0;
    for (long long k = indptr[r]; k < indptr[r + 1]; ++k) {
        const C b = src[indices[k]];
        const double2 v = vals[k];
        ar += v.x * (double)b.x - v.y * (double)b.y;
        ai += v.x * (double)b.y + v.y * (double)b.x;
    }
    const double orr = cr * ar - ci * ai, oi = cr * ai + ci * ar;
    C o;
    if (accumulate) {
        const C d = dst[r];
        o.x = (decltype(o.x))((double)d.x + orr);
        o.y = (decltype(o.y))((double)d.y + oi);
    } else {
        o.x = (decltype(o.x))orr;
        o.y = (decltype(o.y))oi;
    }
    dst[r] = o;
}

// ---- <bra| G_j |ket> for local operators ---------------------------------------------------------------
constexpr int TR_MAX_OPS = 64;   // operators per launch (kernel-parameter bank)
constexpr int TR_CTAS = 148;     // CTAs per operator

struct TransOp {
    int k;          // 1 or 2 bits
    int b0, b1;     // ascending amplitude-index bit positions
    int pad;
    double2 m[16];  // row-major 2^k x 2^k, matrix index bit i <-> bit b_i
};
struct TransParams {
    int nbits, nops;
    int first;    // this launch handles the operators first .. first + gridDim.y - 1
    int pstride;  // partial-sum slots per operator
    int pslot0;   // first slot of this launch
    int pad;
    TransOp op[TR_MAX_OPS];
};

template <typename C>
__global__ void __launch_bounds__(256) transition_kernel(const C* __restrict__ bra, const C* __restrict__ ket,
                                                         const __grid_constant__ TransParams p, double* __restrict__ partials) {
    __shared__ double red[8][2];
    const TransOp& op = p.op[p.first + blockIdx.y];
    const int k = op.k, D = 1 << k;
    const uint64_t ngroups = 1ull << (p.nbits - k);
    double re = 0.0, im = 0.0;
    for (uint64_t g = (uint64_t)blockIdx.x * 256 + threadIdx.x; g < ngroups; g += (uint64_t)gridDim.x * 256) {
        // insert a zero at b0 (and b1)
        uint64_t base = ((g >> op.b0) << (op.b0 + 1)) | (g & ((1ull << op.b0) - 1ull));
        if (k == 2) base = ((base >> op.b1) << (op.b1 + 1)) | (base & ((1ull << op.b1) - 1ull));
        double kr[4], ki[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            if (b < D) {
                const uint64_t idx = base | ((uint64_t)(b & 1) << op.b0) | (k == 2 ? ((uint64_t)(b >> 1) << op.b1) : 0ull);
                const C v = ket[idx];
                kr[b] = (double)v.x;
                ki[b] = (double)v.y;
            }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            if (a < D) {
                double tr = 0.0, ti = 0.0;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    if (b < D) {
                        const double2 m = op.m[a * D + b];
                        tr += m.x * kr[b] - m.y * ki[b];
                        ti += m.x * ki[b] + m.y * kr[b];
                    }
                }
                const uint64_t idx = base | ((uint64_t)(a & 1) << op.b0) | (k == 2 ? ((uint64_t)(a >> 1) << op.b1) : 0ull);
                const C u = bra[idx];
                re += (double)u.x * tr + (double)u.y * ti;
                im += (double)u.x * ti - (double)u.y * tr;
            }
        }
    }
    re = sp_warp_sum(re);
    im = sp_warp_sum(im);
    const int tid = threadIdx.x;
    if ((tid & 31) == 0) {
        red[tid >> 5][0] = re;
        red[tid >> 5][1] = im;
    }
    __syncthreads();
    if (tid == 0) {
        re = im = 0.0;
        for (int w = 0; w < 8; ++w) {
            re += red[w][0];
            im += red[w][1];
        }
        double* o = partials + ((uint64_t)(p.first + blockIdx.y) * p.pstride + p.pslot0 + blockIdx.x) * 2;
        o[0] = re;
        o[1] = im;
    }
}

// ---- the same matrix elements, several operators per load ------------------------------------------
// transition_kernel reads the two states once per operator.  Here the operators of a call are packed into
// BIT GROUPS of four amplitude-index bits: a thread loads the 16 amplitudes of both states that differ in those
// bits once and evaluates every operator of the group (1-bit operators on any of the four bits, 2-bit operators
// on any pair of them) from registers -- one read of the two states per bit group instead of one per operator
// (a layer of 28 single-qubit taps: 7 reads instead of 28).  Products in the state's precision, accumulation in
// float64 per thread (shared-memory slots, no atomics), fixed-order reductions.
constexpr int TT_MAX_GROUPS = 16;  // bit groups per launch (kernel-parameter bank)
constexpr int TT_MAX_OPS = 6;      // operators per bit group

struct TTOp {
    int kind;  // 0..3: 1-bit operator on group position p; 4..9: 2-bit operator on positions (0,1)(0,2)(0,3)(1,2)(1,3)(2,3)
    int out;   // index of the operator in the caller's list
    int pad[2];
    double2 m[16];
};
struct TTGroup {
    int bit[4];  // ascending amplitude-index bits of the group
    int nops;
    int pad[3];
    TTOp op[TT_MAX_OPS];
};
struct TTParams {
    int nbits, ngroups, pstride, pslot0;
    TTGroup g[TT_MAX_GROUPS];
};

template <typename C, typename R, int P0>
__device__ __forceinline__ void tt_one(const C* ps, const C* lm, const double2* m, double& cr, double& ci) {
    const R m00r = (R)m[0].x, m00i = (R)m[0].y, m01r = (R)m[1].x, m01i = (R)m[1].y;
    const R m10r = (R)m[2].x, m10i = (R)m[2].y, m11r = (R)m[3].x, m11i = (R)m[3].y;
    R ar = 0, ai = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int lo = r & ((1 << P0) - 1);
        const int i0 = ((r >> P0) << (P0 + 1)) | lo, i1 = i0 | (1 << P0);
        const R t0r = m00r * ps[i0].x - m00i * ps[i0].y + m01r * ps[i1].x - m01i * ps[i1].y;
        const R t0i = m00r * ps[i0].y + m00i * ps[i0].x + m01r * ps[i1].y + m01i * ps[i1].x;
        const R t1r = m10r * ps[i0].x - m10i * ps[i0].y + m11r * ps[i1].x - m11i * ps[i1].y;
        const R t1i = m10r * ps[i0].y + m10i * ps[i0].x + m11r * ps[i1].y + m11i * ps[i1].x;
        ar += lm[i0].x * t0r + lm[i0].y * t0i + lm[i1].x * t1r + lm[i1].y * t1i;
        ai += lm[i0].x * t0i - lm[i0].y * t0r + lm[i1].x * t1i - lm[i1].y * t1r;
    }
    cr = (double)ar;
    ci = (double)ai;
}

template <typename C, typename R, int P0, int P1>
__device__ __forceinline__ void tt_two(const C* ps, const C* lm, const double2* m, double& cr, double& ci) {
    R ar = 0, ai = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        int base = 0, rb = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b)
            if (b != P0 && b != P1) {
                base |= ((r >> rb) & 1) << b;
                ++rb;
            }
        const int idx[4] = {base, base | (1 << P0), base | (1 << P1), base | (1 << P0) | (1 << P1)};
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            R tr = 0, ti = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const R mr = (R)m[a * 4 + b].x, mi = (R)m[a * 4 + b].y;
                tr += mr * ps[idx[b]].x - mi * ps[idx[b]].y;
                ti += mr * ps[idx[b]].y + mi * ps[idx[b]].x;
            }
            ar += lm[idx[a]].x * tr + lm[idx[a]].y * ti;
            ai += lm[idx[a]].x * ti - lm[idx[a]].y * tr;
        }
    }
    cr = (double)ar;
    ci = (double)ai;
}

template <typename C, typename R>
__global__ void __launch_bounds__(256) transition_tile_kernel(const C* __restrict__ bra, const C* __restrict__ ket,
                                                              const __grid_constant__ TTParams p, double* __restrict__ partials) {
    __shared__ double sacc[TT_MAX_OPS][256][2];
    __shared__ double red[8][2];
    const TTGroup& G = p.g[blockIdx.y];
    const int tid = threadIdx.x;
    for (int o = 0; o < TT_MAX_OPS; ++o) sacc[o][tid][0] = sacc[o][tid][1] = 0.0;
    const int b0 = G.bit[0], b1 = G.bit[1], b2 = G.bit[2], b3 = G.bit[3];
    const uint64_t ngi = 1ull << (p.nbits - 4);
    for (uint64_t gi = (uint64_t)blockIdx.x * 256 + tid; gi < ngi; gi += (uint64_t)gridDim.x * 256) {
        uint64_t base = ((gi >> b0) << (b0 + 1)) | (gi & ((1ull << b0) - 1ull));
        base = ((base >> b1) << (b1 + 1)) | (base & ((1ull << b1) - 1ull));
        base = ((base >> b2) << (b2 + 1)) | (base & ((1ull << b2) - 1ull));
        base = ((base >> b3) << (b3 + 1)) | (base & ((1ull << b3) - 1ull));
        C ps[16], lm[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const uint64_t idx = base | ((uint64_t)(t & 1) << b0) | ((uint64_t)((t >> 1) & 1) << b1) | ((uint64_t)((t >> 2) & 1) << b2) |
                                 ((uint64_t)((t >> 3) & 1) << b3);
            ps[t] = ket[idx];
            lm[t] = bra[idx];
        }
#pragma unroll 1
        for (int o = 0; o < G.nops; ++o) {
            const double2* m = G.op[o].m;
            double cr = 0.0, ci = 0.0;
            switch (G.op[o].kind) {
                case 0: tt_one<C, R, 0>(ps, lm, m, cr, ci); break;
                case 1: tt_one<C, R, 1>(ps, lm, m, cr, ci); break;
                case 2: tt_one<C, R, 2>(ps, lm, m, cr, ci); break;
                case 3: tt_one<C, R, 3>(ps, lm, m, cr, ci); break;
                case 4: tt_two<C, R, 0, 1>(ps, lm, m, cr, ci); break;
                case 5: tt_two<C, R, 0, 2>(ps, lm, m, cr, ci); break;
                case 6: tt_two<C, R, 0, 3>(ps, lm, m, cr, ci); break;
                case 7: tt_two<C, R, 1, 2>(ps, lm, m, cr, ci); break;
                case 8: tt_two<C, R, 1, 3>(ps, lm, m, cr, ci); break;
                default: tt_two<C, R, 2, 3>(ps, lm, m, cr, ci); break;
            }
            sacc[o][tid][0] += cr;
            sacc[o][tid][1] += ci;
        }
    }
    for (int o = 0; o < G.nops; ++o) {
        double re = sp_warp_sum(sacc[o][tid][0]);
        double im = sp_warp_sum(sacc[o][tid][1]);
        __syncthreads();
        if ((tid & 31) == 0) {
            red[tid >> 5][0] = re;
            red[tid >> 5][1] = im;
        }
        __syncthreads();
        if (tid == 0) {
            re = im = 0.0;
            for (int w = 0; w < 8; ++w) {
                re += red[w][0];
                im += red[w][1];
            }
            double* out = partials + ((uint64_t)G.op[o].out * p.pstride + p.pslot0 + blockIdx.x) * 2;
            out[0] = re;
            out[1] = im;
        }
    }
}

// fixed-order sum of the partial slots of every operator; out index through a small map (the host sorts the
// operators into chunk-local and far ones)
struct TransFinalParams {
    int nslots;
    int map[TR_MAX_OPS];
};
__global__ void __launch_bounds__(COO_THREADS) trans_final_kernel(const double* __restrict__ partials, const __grid_constant__ TransFinalParams f,
                                                                   double* __restrict__ out) {
    __shared__ double red[COO_THREADS / 32][2];
    const double* src = partials + (uint64_t)blockIdx.x * f.nslots * 2;
    const int tid = threadIdx.x;
    double re = 0.0, im = 0.0;
    for (int c = tid; c < f.nslots; c += COO_THREADS) {
        re += src[2 * c];
        im += src[2 * c + 1];
    }
    re = sp_warp_sum(re);
    im = sp_warp_sum(im);
    if ((tid & 31) == 0) {
        red[tid >> 5][0] = re;
        red[tid >> 5][1] = im;
    }
    __syncthreads();
    if (tid == 0) {
        re = im = 0.0;
        for (int w = 0; w < COO_THREADS / 32; ++w) {
            re += red[w][0];
            im += red[w][1];
        }
        out[2 * f.map[blockIdx.x]] = re;
        out[2 * f.map[blockIdx.x] + 1] = im;
    }
}


}  // namespace tcb

using namespace tcb;

extern "C" {

size_t tcb200_coo_expectation_workspace_bytes(int64_t nnz, int64_t batch) {
    return (size_t)(batch < 1 ? 1 : batch) * coo_grid(nnz) * 2 * sizeof(double);
}

int tcb200_coo_expectation(const void* state, int nbits, int dtype, int64_t nnz, const int64_t* rows_dev, const int64_t* cols_dev,
                           const void* vals_dev, double* out_dev, int64_t batch, void* workspace, size_t ws_bytes, void* stream) {
    if (!state || !out_dev) return fail(TCB200_ERR_ARG, "NULL argument");
    if (dtype != TCB200_C64 && dtype != TCB200_C128) return fail(TCB200_ERR_ARG, "bad dtype %d", dtype);
    if (nbits < 1 || nbits > 40) return fail(TCB200_ERR_ARG, "nbits=%d out of range", nbits);
    if (nnz < 0) return fail(TCB200_ERR_ARG, "nnz=%lld out of range", (long long)nnz);
    if (nnz > 0 && (!rows_dev || !cols_dev || !vals_dev)) return fail(TCB200_ERR_ARG, "NULL operator array");
    if (batch < 1 || batch > 65535) return fail(TCB200_ERR_ARG, "batch=%lld out of range", (long long)batch);
    if (!workspace || ws_bytes < tcb200_coo_expectation_workspace_bytes(nnz, batch))
        return fail(TCB200_ERR_WORKSPACE, "sparse expectation needs %zu bytes of workspace", tcb200_coo_expectation_workspace_bytes(nnz, batch));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned g = coo_grid(nnz);
    double* partials = static_cast<double*>(workspace);
    const dim3 grid(g, (unsigned)batch);
    if (dtype == TCB200_C64)
        coo_expect_kernel<float2><<<grid, COO_THREADS, 0, st>>>(static_cast<const float2*>(state), nbits, (long long)nnz,
                                                               reinterpret_cast<const long long*>(rows_dev),
                                                               reinterpret_cast<const long long*>(cols_dev),
                                                               static_cast<const double2*>(vals_dev), partials);
    else
        coo_expect_kernel<double2><<<grid, COO_THREADS, 0, st>>>(static_cast<const double2*>(state), nbits, (long long)nnz,
                                                                reinterpret_cast<const long long*>(rows_dev),
                                                                reinterpret_cast<const long long*>(cols_dev),
                                                                static_cast<const double2*>(vals_dev), partials);
    TCB_LAUNCH_CHECK("coo_expect_kernel");
    coo_final_kernel<<<(unsigned)batch, COO_THREADS, 0, st>>>(partials, (int)g, out_dev);
    TCB_LAUNCH_CHECK("coo_final_kernel");
    return 0;
}

size_t tcb200_apply_pauli_sum_workspace_bytes(int nterms) { return (size_t)(nterms < 1 ? 1 : nterms) * sizeof(PauliTerm); }

int tcb200_apply_pauli_sum(const void* src, void* dst, int nbits, int dtype, int nterms, const uint64_t* flip, const uint64_t* sign,
                           const double* coef, int64_t batch, void* workspace, size_t ws_bytes, void* stream) {
    if (!src || !dst || src == dst) return fail(TCB200_ERR_ARG, "src / dst must be two distinct buffers");
    if (dtype != TCB200_C64 && dtype != TCB200_C128) return fail(TCB200_ERR_ARG, "bad dtype %d", dtype);
    if (nbits < 1 || nbits > 40) return fail(TCB200_ERR_ARG, "nbits=%d out of range", nbits);
    if (nterms < 1 || !flip || !sign || !coef) return fail(TCB200_ERR_ARG, "nterms=%d / NULL term arrays", nterms);
    if (batch < 1 || batch > 65535) return fail(TCB200_ERR_ARG, "batch=%lld out of range", (long long)batch);
    if (!workspace || ws_bytes < tcb200_apply_pauli_sum_workspace_bytes(nterms))
        return fail(TCB200_ERR_WORKSPACE, "apply_pauli_sum needs %zu bytes of workspace", tcb200_apply_pauli_sum_workspace_bytes(nterms));
    std::vector<PauliTerm> t((size_t)nterms);
    for (int i = 0; i < nterms; ++i) {
        if (flip[i] >> nbits || sign[i] >> nbits) return fail(TCB200_ERR_ARG, "term %d touches a bit outside the state", i);
        t[i].flip = flip[i];
        t[i].sign = sign[i];
        t[i].cr = coef[2 * i];
        t[i].ci = coef[2 * i + 1];
    }
    std::stable_sort(t.begin(), t.end(), [](const PauliTerm& a, const PauliTerm& b) { return a.flip < b.flip; });
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TCB_CUDA(cudaMemcpyAsync(workspace, t.data(), (size_t)nterms * sizeof(PauliTerm), cudaMemcpyHostToDevice, st));
    TCB_CUDA(cudaStreamSynchronize(st));  // `t` is pageable host memory and goes out of scope
    const uint64_t ctas = ((1ull << nbits) + 255) / 256;
    if (ctas > 0x7fffffffull) return fail(TCB200_ERR_UNSUPPORTED, "state too large for one grid");
    const dim3 grid((unsigned)ctas, (unsigned)batch);
    if (dtype == TCB200_C64)
        pauli_sum_kernel<float2><<<grid, 256, 0, st>>>(static_cast<const float2*>(src), static_cast<float2*>(dst), nbits, nterms,
                                                      static_cast<const PauliTerm*>(workspace));
    else
        pauli_sum_kernel<double2><<<grid, 256, 0, st>>>(static_cast<const double2*>(src), static_cast<double2*>(dst), nbits, nterms,
                                                       static_cast<const PauliTerm*>(workspace));
    TCB_LAUNCH_CHECK("pauli_sum_kernel");
    return 0;
}

int tcb200_csr_matvec(const void* src, void* dst, int nbits, int dtype, const int64_t* indptr_dev, const int64_t* indices_dev,
                      const void* vals_dev, double coef_re, double coef_im, int accumulate, void* stream) {
    if (!src || !dst || src == dst || !indptr_dev) return fail(TCB200_ERR_ARG, "src / dst must be two distinct buffers, indptr non-NULL");
    if (dtype != TCB200_C64 && dtype != TCB200_C128) return fail(TCB200_ERR_ARG, "bad dtype %d", dtype);
    if (nbits < 1 || nbits > 40) return fail(TCB200_ERR_ARG, "nbits=%d out of range", nbits);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const uint64_t ctas = ((1ull << nbits) + 255) / 256;
    if (ctas > 0x7fffffffull) return fail(TCB200_ERR_UNSUPPORTED, "state too large for one grid");
    if (dtype == TCB200_C64)
        csr_matvec_kernel<float2><<<(unsigned)ctas, 256, 0, st>>>(static_cast<const float2*>(src), static_cast<float2*>(dst), nbits,
                                                                  reinterpret_cast<const long long*>(indptr_dev),
                                                                  reinterpret_cast<const long long*>(indices_dev),
                                                                  static_cast<const double2*>(vals_dev), coef_re, coef_im, accumulate);
    else
        csr_matvec_kernel<double2><<<(unsigned)ctas, 256, 0, st>>>(static_cast<const double2*>(src), static_cast<double2*>(dst), nbits,
                                                                   reinterpret_cast<const long long*>(indptr_dev),
                                                                   reinterpret_cast<const long long*>(indices_dev),
                                                                   static_cast<const double2*>(vals_dev), coef_re, coef_im, accumulate);
    TCB_LAUNCH_CHECK("csr_matvec_kernel");
    return 0;
}

int tcb200_transition_local_max_ops(void) { return TR_MAX_OPS; }

// Operators whose bits all lie below the chunk size are evaluated chunk by chunk: every launch covers one
// L2-sized chunk of both states for ALL such operators, so the chunk comes from HBM once and the other
// operators read it from L2 (one read of the two states for the whole group instead of one per operator).
// Operators with a bit above the chunk size stream the whole states, one read each.
static int trans_chunk_bits(int dtype) {
    const char* e = getenv("TCB200_TRANS_CHUNK_BITS");  // tests force tiny chunks
    if (e && e[0]) {
        const int v = atoi(e);
        if (v >= 2) return v;
    }
    return dtype == TCB200_C64 ? 22 : 21;  // 2 x 32 MiB of amplitudes per chunk: half of the 126 MB L2
}
static int trans_nchunks(int nbits, int dtype) {
    const int cb = trans_chunk_bits(dtype);
    return nbits >= cb + 2 ? 1 << (nbits - cb) : 1;
}

size_t tcb200_transition_local_workspace_bytes(int nops, int nbits, int dtype) {
    return (size_t)(nops < 1 ? 1 : nops) * TR_CTAS * (size_t)trans_nchunks(nbits, dtype) * 2 * sizeof(double);
}

int tcb200_transition_local(const void* bra, const void* ket, int nbits, int dtype, int nops, const int* ops_k, const int* ops_bits,
                            const double* ops_mats, double* out_dev, void* workspace, size_t ws_bytes, void* stream) {
    if (!bra || !ket || !out_dev || !ops_k || !ops_bits || !ops_mats) return fail(TCB200_ERR_ARG, "NULL argument");
    if (dtype != TCB200_C64 && dtype != TCB200_C128) return fail(TCB200_ERR_ARG, "bad dtype %d", dtype);
    if (nbits < 1 || nbits > 40) return fail(TCB200_ERR_ARG, "nbits=%d out of range", nbits);
    if (nops < 1 || nops > TR_MAX_OPS) return fail(TCB200_ERR_ARG, "nops=%d out of range (max %d per call)", nops, TR_MAX_OPS);
    if (!workspace || ws_bytes < tcb200_transition_local_workspace_bytes(nops, nbits, dtype))
        return fail(TCB200_ERR_WORKSPACE, "transition_local needs %zu bytes of workspace", tcb200_transition_local_workspace_bytes(nops, nbits, dtype));
    static thread_local TransParams* tp = nullptr;
    if (!tp) tp = new TransParams();
    TransParams& p = *tp;
    TransFinalParams fin;
    const int cb = trans_chunk_bits(dtype);
    int nchunks = trans_nchunks(nbits, dtype);
    // validate, then order the operators: chunk-local ones first
    struct In { int k, b0, b1; const double* m; int idx; };
    In in[TR_MAX_OPS];
    const int* b = ops_bits;
    const double* m = ops_mats;
    int nlocal = 0;
    for (int o = 0; o < nops; ++o) {
        const int k = ops_k[o];
        if (k < 1 || k > 2) return fail(TCB200_ERR_UNSUPPORTED, "local operator of %d bits (max 2)", k);
        if (k > nbits) return fail(TCB200_ERR_ARG, "operator wider than the state");
        for (int i = 0; i < k; ++i) {
            if (b[i] < 0 || b[i] >= nbits) return fail(TCB200_ERR_ARG, "bit %d out of range", b[i]);
            if (i > 0 && b[i] <= b[i - 1]) return fail(TCB200_ERR_ARG, "bits must be strictly ascending");
        }
        in[o] = In{k, b[0], k == 2 ? b[1] : 0, m, o};
        if (b[k - 1] < cb) ++nlocal;
        b += k;
        m += 2 * (1 << (2 * k));
    }
    {
        const char* e = getenv("TCB200_TRANS_TILE");  // =0: one read of the states per operator (transition_kernel)
        if (nbits >= 4 && !(e && e[0] == '0')) {
            // ---- bit groups: operators sharing <= 4 bits are evaluated from one load of 16 + 16 amplitudes ----
            struct Grp { int bits[4]; int nb; int ops[TT_MAX_OPS]; int nops; };
            std::vector<Grp> grps;
            std::vector<int> order(nops);
            for (int o = 0; o < nops; ++o) order[o] = o;
            std::stable_sort(order.begin(), order.end(), [&](int a, int c) { return in[a].b0 < in[c].b0; });
            for (int oi : order) {
                const In& op = in[oi];
                const int ob[2] = {op.b0, op.b1};
                int placed = -1;
                for (int gi = (int)grps.size() - 1; gi >= 0 && gi >= (int)grps.size() - 4 && placed < 0; --gi) {
                    Grp& g = grps[gi];
                    if (g.nops >= TT_MAX_OPS) continue;
                    int add = 0;
                    for (int i = 0; i < op.k; ++i) {
                        bool have = false;
                        for (int j = 0; j < g.nb; ++j) have = have || g.bits[j] == ob[i];
                        if (!have) ++add;
                    }
                    if (g.nb + add > 4) continue;
                    for (int i = 0; i < op.k; ++i) {
                        bool have = false;
                        for (int j = 0; j < g.nb; ++j) have = have || g.bits[j] == ob[i];
                        if (!have) g.bits[g.nb++] = ob[i];
                    }
                    g.ops[g.nops++] = oi;
                    placed = gi;
                }
                if (placed < 0) {
                    Grp g;
                    g.nb = 0;
                    g.nops = 0;
                    for (int i = 0; i < op.k; ++i) g.bits[g.nb++] = ob[i];
                    g.ops[g.nops++] = oi;
                    grps.push_back(g);
                }
            }
            static thread_local TTParams* ttp = nullptr;
            if (!ttp) ttp = new TTParams();
            TTParams& q = *ttp;
            const size_t amp = dtype == TCB200_C64 ? 8 : 16;
            std::vector<int> local_ids, far_ids;
            for (size_t gi = 0; gi < grps.size(); ++gi) {
                Grp& g = grps[gi];
                for (int cand = 0; g.nb < 4; ++cand) {  // pad with the lowest free bits: the tile is always 16 amplitudes
                    bool have = false;
                    for (int j = 0; j < g.nb; ++j) have = have || g.bits[j] == cand;
                    if (!have) g.bits[g.nb++] = cand;
                }
                std::sort(g.bits, g.bits + 4);
                (nchunks > 1 && g.bits[3] < cb ? local_ids : far_ids).push_back((int)gi);
            }
            if (local_ids.size() < 2) {  // nothing to share inside a chunk
                far_ids.insert(far_ids.end(), local_ids.begin(), local_ids.end());
                local_ids.clear();
                nchunks = 1;
            }
            cudaStream_t st = static_cast<cudaStream_t>(stream);
            double* partials = static_cast<double*>(workspace);
            const int pstride = TR_CTAS * nchunks;
            if (nchunks > 1) TCB_CUDA(cudaMemsetAsync(partials, 0, (size_t)nops * pstride * 2 * sizeof(double), st));
            static const int pair_index[4][4] = {{-1, 4, 5, 6}, {-1, -1, 7, 8}, {-1, -1, -1, 9}, {-1, -1, -1, -1}};
            auto fill = [&](const std::vector<int>& ids, size_t from, int count) {
                q.ngroups = count;
                for (int c = 0; c < count; ++c) {
                    const Grp& g = grps[ids[from + c]];
                    TTGroup& tg = q.g[c];
                    for (int j = 0; j < 4; ++j) tg.bit[j] = g.bits[j];
                    tg.nops = g.nops;
                    for (int o = 0; o < g.nops; ++o) {
                        const In& op = in[g.ops[o]];
                        int pos[2] = {0, 0};
                        const int ob[2] = {op.b0, op.b1};
                        for (int i = 0; i < op.k; ++i)
                            for (int j = 0; j < 4; ++j)
                                if (g.bits[j] == ob[i]) pos[i] = j;
                        tg.op[o].kind = op.k == 1 ? pos[0] : pair_index[pos[0]][pos[1]];
                        tg.op[o].out = op.idx;
                        const int D = 1 << op.k;
                        for (int i = 0; i < D * D; ++i) tg.op[o].m[i] = make_double2(op.m[2 * i], op.m[2 * i + 1]);
                    }
                }
            };
            auto launch_tiles = [&](int nb, size_t chunk, int slot0) {
                q.nbits = nb;
                q.pstride = pstride;
                q.pslot0 = slot0;
                const unsigned char* br = static_cast<const unsigned char*>(bra) + (chunk << nb) * amp;
                const unsigned char* kt = static_cast<const unsigned char*>(ket) + (chunk << nb) * amp;
                const dim3 grid(TR_CTAS, (unsigned)q.ngroups);
                if (dtype == TCB200_C64)
                    transition_tile_kernel<float2, float><<<grid, 256, 0, st>>>(reinterpret_cast<const float2*>(br), reinterpret_cast<const float2*>(kt), q, partials);
                else
                    transition_tile_kernel<double2, double><<<grid, 256, 0, st>>>(reinterpret_cast<const double2*>(br), reinterpret_cast<const double2*>(kt), q, partials);
            };
            for (size_t from = 0; from < local_ids.size(); from += TT_MAX_GROUPS) {
                fill(local_ids, from, (int)std::min<size_t>(TT_MAX_GROUPS, local_ids.size() - from));
                for (int c = 0; c < nchunks; ++c) {
                    launch_tiles(cb, (size_t)c, c * TR_CTAS);
                    TCB_LAUNCH_CHECK("transition_tile_kernel");
                }
            }
            for (size_t from = 0; from < far_ids.size(); from += TT_MAX_GROUPS) {
                fill(far_ids, from, (int)std::min<size_t>(TT_MAX_GROUPS, far_ids.size() - from));
                launch_tiles(nbits, 0, 0);
                TCB_LAUNCH_CHECK("transition_tile_kernel");
            }
            fin.nslots = pstride;
            for (int o = 0; o < nops; ++o) fin.map[o] = o;
            trans_final_kernel<<<(unsigned)nops, COO_THREADS, 0, st>>>(partials, fin, out_dev);
            TCB_LAUNCH_CHECK("trans_final_kernel");
            return 0;
        }
    }
    if (nlocal < 2) nchunks = 1;  // nothing to share
    int lo = 0, hi = nchunks > 1 ? nlocal : 0;
    for (int o = 0; o < nops; ++o) {
        const bool local = nchunks > 1 && (in[o].k == 2 ? in[o].b1 : in[o].b0) < cb;
        const int at = nchunks > 1 ? (local ? lo++ : hi++) : o;
        p.op[at].k = in[o].k;
        p.op[at].b0 = in[o].b0;
        p.op[at].b1 = in[o].b1;
        const int D = 1 << in[o].k;
        for (int i = 0; i < D * D; ++i) p.op[at].m[i] = make_double2(in[o].m[2 * i], in[o].m[2 * i + 1]);
        fin.map[at] = in[o].idx;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double* partials = static_cast<double*>(workspace);
    const int pstride = TR_CTAS * nchunks;
    p.nops = nops;
    p.pstride = pstride;
    fin.nslots = pstride;
    if (nchunks > 1) TCB_CUDA(cudaMemsetAsync(partials, 0, (size_t)nops * pstride * 2 * sizeof(double), st));
    const size_t amp = dtype == TCB200_C64 ? 8 : 16;
    auto launch = [&](int first, int count, int nb, size_t chunk, int slot0) {
        p.nbits = nb;
        p.first = first;
        p.pslot0 = slot0;
        const unsigned char* br = static_cast<const unsigned char*>(bra) + (chunk << nb) * amp;
        const unsigned char* kt = static_cast<const unsigned char*>(ket) + (chunk << nb) * amp;
        const dim3 grid(TR_CTAS, (unsigned)count);
        if (dtype == TCB200_C64)
            transition_kernel<float2><<<grid, 256, 0, st>>>(reinterpret_cast<const float2*>(br), reinterpret_cast<const float2*>(kt), p, partials);
        else
            transition_kernel<double2><<<grid, 256, 0, st>>>(reinterpret_cast<const double2*>(br), reinterpret_cast<const double2*>(kt), p, partials);
    };
    if (nchunks > 1) {
        for (int c = 0; c < nchunks; ++c) {
            launch(0, nlocal, cb, (size_t)c, c * TR_CTAS);
            TCB_LAUNCH_CHECK("transition_kernel");
        }
        if (nops > nlocal) {
            launch(nlocal, nops - nlocal, nbits, 0, 0);
            TCB_LAUNCH_CHECK("transition_kernel");
        }
    } else {
        launch(0, nops, nbits, 0, 0);
        TCB_LAUNCH_CHECK("transition_kernel");
    }
    trans_final_kernel<<<(unsigned)nops, COO_THREADS, 0, st>>>(partials, fin, out_dev);
    TCB_LAUNCH_CHECK("trans_final_kernel");
    return 0;
}

}  // extern "C"
