// Pauli-string expectation values: up to TCB200_MAX_TERMS strings per read of the state.
//
// A CTA walks over tiles (same gather geometry as the apply kernels, linear layout).  For a
// term with local flip mask f and local sign mask s the contribution of tile element e is
//     conj(psi_e) * psi_{e^f} * (-1)^{popc(e & s)}     (quantum.py:1461-1482)
// times a per-tile sign from the sign bits outside the tile.  Per-thread partials of one tile
// are formed in the state's real type (16 products), then accumulated in float64 across
// tiles; warp-shuffle + block reduce at the end; the cross-CTA sum and the (-i)^{n_y} phase
// are applied by a second tiny kernel in a fixed order, so results are deterministic.
#include <string.h>

#include "common.cuh"

namespace tcb {

struct ExpectParams {
    const void* state;
    double* partials;  // [batch][gridDim.x][MAXT][2]
    TileGeom g;
    int nterms;
    uint32_t flip_l[TCB200_MAX_TERMS];
    uint32_t sign_l[TCB200_MAX_TERMS];
    uint64_t sign_hi[TCB200_MAX_TERMS];  // sign bits outside the tile (global positions)
    uint32_t odd_mask;                   // bit t: the string holds an odd number of Y's
};

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename Real>
__global__ void __launch_bounds__(256, 3) expect_kernel(const __grid_constant__ ExpectParams p) {
    using C = typename CT<Real>::type;
    constexpr int MT = TCB200_MAX_TERMS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    C* tile = reinterpret_cast<C*>(smem_raw);
    __shared__ uint64_t rowoff[256];
    __shared__ double red[8][MT];

    const int tid = threadIdx.x;
    const int nthr = blockDim.x;
    if (tid < (1 << p.g.h)) rowoff[tid] = row_offset(p.g, tid);
    __syncthreads();
    const C* vec = static_cast<const C*>(p.state) + ((uint64_t)blockIdx.y << p.g.n);
    const uint64_t ntiles = 1ull << (p.g.n - p.g.T);
    const uint32_t tsz = 1u << p.g.T;

    double acc[MT];  // the non-zero component of every term (expect_tile_terms)
#pragma unroll
    for (int t = 0; t < MT; ++t) acc[t] = 0.0;

    for (uint64_t tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
        const uint64_t base = tile_base(p.g, tl);
        stage_in<C, false>(p.g, vec, base, tile, rowoff, tid, nthr);
        cp_async_wait_all();
        __syncthreads();
        {
            Real pv[MT];
            expect_tile_terms<C, Real, MT>(tile, tsz, p.nterms, p.flip_l, p.sign_l, p.odd_mask, tid, nthr, pv);
#pragma unroll
            for (int t = 0; t < MT; ++t) {
                if (t < p.nterms) {
                    const bool neg = parity64(base & p.sign_hi[t]);
                    acc[t] += neg ? -(double)pv[t] : (double)pv[t];
                }
            }
        }
        __syncthreads();  // tile is overwritten by the next iteration
    }
    // CTA reduction
    const int w = tid >> 5, l = tid & 31;
#pragma unroll
    for (int t = 0; t < MT; ++t) {
        if (t < p.nterms) {
            const double r = warp_sum_d(acc[t]);
            if (l == 0) red[w][t] = r;
        }
    }
    __syncthreads();
    if (tid < p.nterms) {
        const int t = tid;
        double s = 0.0;
        const int nw = nthr >> 5;
        for (int ww = 0; ww < nw; ++ww) s += red[ww][t];
        const bool odd = (p.odd_mask >> t) & 1u;
        double* o = p.partials + (((uint64_t)blockIdx.y * gridDim.x + blockIdx.x) * MT + t) * 2;
        o[0] = odd ? 0.0 : s;
        o[1] = odd ? s : 0.0;
    }
}

struct ExpectFinalParams {
    const double* partials;
    double* out;  // [batch][nterms][2]
    int nctas;
    int nterms;
    int ny[TCB200_MAX_TERMS];
};

// grid (batch, nterms), 256 threads: strided partial sums, then a fixed shuffle / shared-memory tree
// (the order of the additions does not depend on timing: results are reproducible bit for bit)
__global__ void __launch_bounds__(256) expect_final_kernel(const __grid_constant__ ExpectFinalParams p) {
    constexpr int MT = TCB200_MAX_TERMS;
    __shared__ double red[8][2];
    const int t = blockIdx.y, tid = threadIdx.x;
    const double* src = p.partials + (uint64_t)blockIdx.x * p.nctas * MT * 2;
    double re = 0.0, im = 0.0;
    for (int c = tid; c < p.nctas; c += 256) {
        re += src[((uint64_t)c * MT + t) * 2];
        im += src[((uint64_t)c * MT + t) * 2 + 1];
    }
    re = warp_sum_d(re);
    im = warp_sum_d(im);
    if ((tid & 31) == 0) {
        red[tid >> 5][0] = re;
        red[tid >> 5][1] = im;
    }
    __syncthreads();
    if (tid != 0) return;
    re = im = 0.0;
    for (int w = 0; w < 8; ++w) {
        re += red[w][0];
        im += red[w][1];
    }
    // times (-i)^ny
    double ore = re, oim = im;
    switch (p.ny[t] & 3) {
        case 1: ore = im; oim = -re; break;
        case 2: ore = -re; oim = -im; break;
        case 3: ore = -im; oim = re; break;
        default: break;
    }
    double* o = p.out + ((uint64_t)blockIdx.x * p.nterms + t) * 2;
    o[0] = ore;
    o[1] = oim;
}

static unsigned expect_grid_x(int nbits, int dtype, int64_t batch) {
    const int T = expect_tile_bits(dtype);
    const uint64_t ntiles = nbits > T ? (1ull << (nbits - T)) : 1ull;
    uint64_t cap = (148ull * 16) / (uint64_t)batch;
    if (cap < 4) cap = 4;
    return (unsigned)(ntiles < cap ? ntiles : cap);
}

}  // namespace tcb

using namespace tcb;

extern "C" {

int tcb200_expect_tile_bits(int dtype) { return expect_tile_bits(dtype); }

size_t tcb200_expect_workspace_bytes(int nbits, int64_t batch) {
    // sized for either dtype
    const unsigned g0 = expect_grid_x(nbits, TCB200_C64, batch), g1 = expect_grid_x(nbits, TCB200_C128, batch);
    const unsigned gx = g0 > g1 ? g0 : g1;
    return (size_t)batch * gx * TCB200_MAX_TERMS * 2 * sizeof(double) + 256;
}

int tcb200_expect_pauli(const void* state, int nbits, int dtype, int nterms,
                        const uint64_t* flip, const uint64_t* sign, const int* ny, int n_hi,
                        const int* tile_hi, double* out_dev, int64_t batch, void* workspace,
                        size_t ws_bytes, void* stream) {
    if (!state || !flip || !sign || !ny || !out_dev || !workspace) return fail(TCB200_ERR_ARG, "NULL argument");
    if (dtype != TCB200_C64 && dtype != TCB200_C128) return fail(TCB200_ERR_ARG, "bad dtype %d", dtype);
    if (nbits < 1 || nbits > 40) return fail(TCB200_ERR_ARG, "nbits=%d out of range", nbits);
    if (nterms < 1 || nterms > TCB200_MAX_TERMS) return fail(TCB200_ERR_ARG, "nterms=%d out of range", nterms);
    if (batch < 1 || batch > 65535) return fail(TCB200_ERR_ARG, "batch=%lld out of range", (long long)batch);
    if (ws_bytes < tcb200_expect_workspace_bytes(nbits, batch)) return fail(TCB200_ERR_WORKSPACE, "workspace too small");
    ExpectParams p;
    memset(&p, 0, sizeof(p));
    const int T = expect_tile_bits(dtype);
    int rc = make_geom_hi(nbits, T, nbits <= T ? 0 : n_hi, tile_hi, &p.g);
    if (rc) return rc;
    p.state = state;
    p.partials = static_cast<double*>(workspace);
    p.nterms = nterms;
    const uint64_t full = nbits >= 64 ? ~0ull : ((1ull << nbits) - 1ull);
    for (int t = 0; t < nterms; ++t) {
        if ((flip[t] & ~full) || (sign[t] & ~full)) return fail(TCB200_ERR_ARG, "mask of term %d exceeds the state", t);
        uint32_t fl = 0, sl = 0;
        uint64_t shi = sign[t];
        for (int b = 0; b < nbits; ++b) {
            const int lb = local_bit(p.g, b);
            if ((flip[t] >> b) & 1ull) {
                if (lb < 0) return fail(TCB200_ERR_ARG, "flip bit %d of term %d is not inside the tile", b, t);
                fl |= 1u << lb;
            }
            if (((sign[t] >> b) & 1ull) && lb >= 0) {
                sl |= 1u << lb;
                shi &= ~(1ull << b);
            }
        }
        p.flip_l[t] = fl;
        p.sign_l[t] = sl;
        p.sign_hi[t] = shi;
        if (ny[t] & 1) p.odd_mask |= 1u << t;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned gx = expect_grid_x(nbits, dtype, batch);
    const size_t esz = dtype == TCB200_C64 ? 8 : 16;
    size_t smem = esz << p.g.T;
    if (smem < 16) smem = 16;
    int tb = p.g.T - (dtype == TCB200_C64 ? 1 : 0);
    if (tb > 8) tb = 8;
    if (tb < 5) tb = 5;
    dim3 grid(gx, (unsigned)batch);
    if (dtype == TCB200_C64)
        expect_kernel<float><<<grid, 1u << tb, smem, st>>>(p);
    else
        expect_kernel<double><<<grid, 1u << tb, smem, st>>>(p);
    TCB_LAUNCH_CHECK("expect_kernel");
    ExpectFinalParams f;
    memset(&f, 0, sizeof(f));
    f.partials = p.partials;
    f.out = out_dev;
    f.nctas = (int)gx;
    f.nterms = nterms;
    for (int t = 0; t < nterms; ++t) f.ny[t] = ny[t];
    expect_final_kernel<<<dim3((unsigned)batch, (unsigned)nterms), 256, 0, st>>>(f);
    TCB_LAUNCH_CHECK("expect_final_kernel");
    return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// Diagonal strings (only I / Z): <P_t> = sum_e |psi_e|^2 (-1)^{popc(e & mask_t)}
// ------------------------------------------------------------------------------------------------
// No partner amplitude is needed, so there is no tile and no shared memory: every thread
// streams 16-byte loads, and the sign of (amplitude, term) is split as
//     popc(e & m) = popc(chunk_base & m) + popc(thread_offset & m) + popc(iteration_offset & m)   (mod 2)
// The iteration part is the same for all threads: a +-1 table in the kernel-parameter constant
// bank, so one (amplitude, term) pair costs exactly ONE FFMA with a constant operand.  The chunk
// part is applied once per chunk (2^13 amplitudes), the thread part once at the end.  32 strings
// (16 for complex128) ride on one read of the state: a QAOA / Ising cost function is one pass.
namespace tcb {

constexpr int ZK = 16;  // 16-byte loads per thread and chunk

template <typename Real>
struct ZCfg {
    static constexpr int ZT = sizeof(Real) == 4 ? 32 : 16;  // strings per launch
    static constexpr int APU = CT<Real>::APU;
    static constexpr int AB = APU == 2 ? 1 : 0;
    static constexpr int CHUNK_BITS = 4 + 8 + AB;  // ZK x 256 threads x APU amplitudes
};

template <typename Real>
struct ZExpectParams {
    const void* state;
    double* partials;  // [batch][gridDim.x][ZT]
    int nbits;
    uint64_t mask[ZCfg<Real>::ZT];
    Real s[ZK][ZCfg<Real>::APU][ZCfg<Real>::ZT];  // (-1)^{popc(iteration offset & mask)}
};

// one chunk of one thread: tmp[t] += sum over its ZK * APU amplitudes of s * |psi|^2
template <typename Real>
TCB_HD void zexpect_chunk(const typename CT<Real>::type* vec, uint64_t base, int tid,
                          const Real (*s)[ZCfg<Real>::APU][ZCfg<Real>::ZT], Real* tmp) {
    using C = typename CT<Real>::type;
    constexpr int ZT = ZCfg<Real>::ZT, APU = ZCfg<Real>::APU;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        Unit16 q[ZK / 2];
#pragma unroll
        for (int i = 0; i < ZK / 2; ++i) {
            const C* src = vec + base + (uint64_t)(half * (ZK / 2) + i) * (256 * APU) + (uint64_t)tid * APU;
#if defined(__CUDA_ARCH__)
            const uint4 r = __ldg(reinterpret_cast<const uint4*>(src));
            q[i].w[0] = r.x; q[i].w[1] = r.y; q[i].w[2] = r.z; q[i].w[3] = r.w;
#else
            q[i] = *reinterpret_cast<const Unit16*>(src);
#endif
        }
#pragma unroll
        for (int i = 0; i < ZK / 2; ++i) {
            const C* a = reinterpret_cast<const C*>(&q[i]);
#pragma unroll
            for (int j = 0; j < APU; ++j) {
                const Real p = a[j].x * a[j].x + a[j].y * a[j].y;
#pragma unroll
                for (int t = 0; t < ZT; ++t) tmp[t] = fma(p, s[half * (ZK / 2) + i][j][t], tmp[t]);
            }
        }
    }
}

template <typename Real>
__global__ void __launch_bounds__(256, 2) zexpect_kernel(const __grid_constant__ ZExpectParams<Real> p) {
    using C = typename CT<Real>::type;
    constexpr int ZT = ZCfg<Real>::ZT, AB = ZCfg<Real>::AB, CB = ZCfg<Real>::CHUNK_BITS;
    __shared__ double red[8][ZT];
    const int tid = threadIdx.x;
    const C* vec = static_cast<const C*>(p.state) + ((uint64_t)blockIdx.y << p.nbits);
    const uint64_t nchunks = 1ull << (p.nbits - CB);
    Real acc[ZT];
#pragma unroll
    for (int t = 0; t < ZT; ++t) acc[t] = 0;
    for (uint64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const uint64_t base = c << CB;
        Real tmp[ZT];
#pragma unroll
        for (int t = 0; t < ZT; ++t) tmp[t] = 0;
        zexpect_chunk<Real>(vec, base, tid, p.s, tmp);
#pragma unroll
        for (int t = 0; t < ZT; ++t) acc[t] += parity64(base & p.mask[t]) ? -tmp[t] : tmp[t];
    }
    const int w = tid >> 5, l = tid & 31;
#pragma unroll
    for (int t = 0; t < ZT; ++t) {
        double v = (double)acc[t];
        if (parity64(((uint64_t)tid << AB) & p.mask[t])) v = -v;
        v = warp_sum_d(v);
        if (l == 0) red[w][t] = v;
    }
    __syncthreads();
    if (tid < ZT) {
        double sum = 0.0;
        for (int ww = 0; ww < 8; ++ww) sum += red[ww][tid];
        p.partials[((uint64_t)blockIdx.y * gridDim.x + blockIdx.x) * ZT + tid] = sum;
    }
}

struct ZFinalParams {
    const double* partials;
    double* out;  // [batch][nterms][2]
    int nctas;
    int nterms;
    int zt;
};

__global__ void __launch_bounds__(256) zexpect_final_kernel(const __grid_constant__ ZFinalParams p) {
    __shared__ double red[8];
    const int t = blockIdx.y, tid = threadIdx.x;
    const double* src = p.partials + (uint64_t)blockIdx.x * p.nctas * p.zt;
    double re = 0.0;
    for (int c = tid; c < p.nctas; c += 256) re += src[(uint64_t)c * p.zt + t];
    re = warp_sum_d(re);
    if ((tid & 31) == 0) red[tid >> 5] = re;
    __syncthreads();
    if (tid != 0) return;
    re = 0.0;
    for (int w = 0; w < 8; ++w) re += red[w];
    double* o = p.out + ((uint64_t)blockIdx.x * p.nterms + t) * 2;
    o[0] = re;
    o[1] = 0.0;
}

static unsigned zexpect_grid_x(int nbits, int chunk_bits, int64_t batch) {
    const uint64_t nchunks = 1ull << (nbits - chunk_bits);
    uint64_t cap = (148ull * 8) / (uint64_t)batch;
    if (cap < 4) cap = 4;
    return (unsigned)(nchunks < cap ? nchunks : cap);
}

template <typename Real>
static void zexpect_fill(ZExpectParams<Real>& p, int nterms, const uint64_t* sign) {
    constexpr int ZT = ZCfg<Real>::ZT, APU = ZCfg<Real>::APU, AB = ZCfg<Real>::AB;
    for (int t = 0; t < ZT; ++t) {
        p.mask[t] = t < nterms ? sign[t] : 0ull;
        for (int k = 0; k < ZK; ++k)
            for (int j = 0; j < APU; ++j) {
                const uint64_t off = ((uint64_t)k << (8 + AB)) | (uint64_t)j;
                p.s[k][j][t] = parity64(off & p.mask[t]) ? (Real)-1 : (Real)1;
            }
    }
}

template <typename Real>
static int launch_zexpect(const void* state, int nbits, int nterms, const uint64_t* sign, double* out_dev,
                          int64_t batch, void* workspace, cudaStream_t st) {
    constexpr int ZT = ZCfg<Real>::ZT, CB = ZCfg<Real>::CHUNK_BITS;
    static thread_local ZExpectParams<Real>* tp = nullptr;
    if (!tp) tp = new ZExpectParams<Real>();
    ZExpectParams<Real>& p = *tp;
    p.state = state;
    p.partials = static_cast<double*>(workspace);
    p.nbits = nbits;
    zexpect_fill<Real>(p, nterms, sign);
    const unsigned gx = zexpect_grid_x(nbits, CB, batch);
    dim3 grid(gx, (unsigned)batch);
    zexpect_kernel<Real><<<grid, 256, 0, st>>>(p);
    TCB_LAUNCH_CHECK("zexpect_kernel");
    ZFinalParams f;
    f.partials = p.partials;
    f.out = out_dev;
    f.nctas = (int)gx;
    f.nterms = nterms;
    f.zt = ZT;
    zexpect_final_kernel<<<dim3((unsigned)batch, (unsigned)nterms), 256, 0, st>>>(f);
    TCB_LAUNCH_CHECK("zexpect_final_kernel");
    return 0;
}

#ifdef TCB200_EMU
// tests/emu only: the same parameter block and per-thread chunk body on the CPU
template <typename Real>
static int emu_zexpect_t(const void* state, int nbits, int nterms, const uint64_t* sign, double* out) {
    using C = typename CT<Real>::type;
    constexpr int ZT = ZCfg<Real>::ZT, AB = ZCfg<Real>::AB, CB = ZCfg<Real>::CHUNK_BITS;
    if (nbits < CB || nterms < 1 || nterms > ZT) return 1;
    ZExpectParams<Real>* pp = new ZExpectParams<Real>();
    ZExpectParams<Real>& p = *pp;
    zexpect_fill<Real>(p, nterms, sign);
    const C* vec = static_cast<const C*>(state);
    double tot[ZT] = {0};
    for (uint64_t c = 0; c < (1ull << (nbits - CB)); ++c)
        for (int tid = 0; tid < 256; ++tid) {
            Real tmp[ZT];
            for (int t = 0; t < ZT; ++t) tmp[t] = 0;
            zexpect_chunk<Real>(vec, c << CB, tid, p.s, tmp);
            for (int t = 0; t < ZT; ++t) {
                double v = parity64((c << CB) & p.mask[t]) ? -(double)tmp[t] : (double)tmp[t];
                if (parity64(((uint64_t)tid << AB) & p.mask[t])) v = -v;
                tot[t] += v;
            }
        }
    for (int t = 0; t < nterms; ++t) out[t] = tot[t];
    delete pp;
    return 0;
}

extern "C" __attribute__((visibility("default"))) int emu_expect_z(const void* state, int nbits, int dtype, int nterms,
                                                                    const uint64_t* sign, double* out) {
    return dtype == TCB200_C64 ? emu_zexpect_t<float>(state, nbits, nterms, sign, out)
                               : emu_zexpect_t<double>(state, nbits, nterms, sign, out);
}
#endif

}  // namespace tcb

extern "C" {

int tcb200_expect_z_max_terms(int dtype) { return dtype == TCB200_C64 ? ZCfg<float>::ZT : ZCfg<double>::ZT; }

int tcb200_expect_z_min_bits(int dtype) { return dtype == TCB200_C64 ? ZCfg<float>::CHUNK_BITS : ZCfg<double>::CHUNK_BITS; }

size_t tcb200_expect_z_workspace_bytes(int nbits, int64_t batch) {
    (void)nbits;
    if (batch < 1) batch = 1;
    // partials [batch][grid.x][32]; grid.x * batch <= 148 * 8 + 4 * batch (zexpect_grid_x)
    return ((size_t)148 * 8 + 4 * (size_t)batch) * 32 * sizeof(double) + 256;
}

int tcb200_expect_z(const void* state, int nbits, int dtype, int nterms, const uint64_t* sign,
                    double* out_dev, int64_t batch, void* workspace, size_t ws_bytes, void* stream) {
    if (!state || !sign || !out_dev || !workspace) return fail(TCB200_ERR_ARG, "NULL argument");
    if (dtype != TCB200_C64 && dtype != TCB200_C128) return fail(TCB200_ERR_ARG, "bad dtype %d", dtype);
    if (nbits < 1 || nbits > 40) return fail(TCB200_ERR_ARG, "nbits=%d out of range", nbits);
    if (nterms < 1 || nterms > tcb200_expect_z_max_terms(dtype)) return fail(TCB200_ERR_ARG, "nterms=%d out of range", nterms);
    if (batch < 1 || batch > 65535) return fail(TCB200_ERR_ARG, "batch=%lld out of range", (long long)batch);
    if (ws_bytes < tcb200_expect_z_workspace_bytes(nbits, batch)) return fail(TCB200_ERR_WORKSPACE, "workspace too small");
    const uint64_t full = (1ull << nbits) - 1ull;
    for (int t = 0; t < nterms; ++t)
        if (sign[t] & ~full) return fail(TCB200_ERR_ARG, "mask of term %d exceeds the state", t);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int cb = dtype == TCB200_C64 ? ZCfg<float>::CHUNK_BITS : ZCfg<double>::CHUNK_BITS;
    if (nbits < cb)
        return fail(TCB200_ERR_UNSUPPORTED, "tcb200_expect_z needs a state of >= %d bits (use tcb200_expect_pauli)", cb);
    if ((size_t)batch * zexpect_grid_x(nbits, cb, batch) * 32 * sizeof(double) > ws_bytes)
        return fail(TCB200_ERR_WORKSPACE, "workspace too small");
    if (dtype == TCB200_C64) return launch_zexpect<float>(state, nbits, nterms, sign, out_dev, batch, workspace, st);
    return launch_zexpect<double>(state, nbits, nterms, sign, out_dev, batch, workspace, st);
}

}  // extern "C"
