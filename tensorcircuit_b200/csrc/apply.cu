// In-place gate application kernels (sm_100a).
//
//   dense_kernel<Real, K, BATCHED>
//       one fused 2^K x 2^K block per launch.  A CTA stages one tile (32 KiB) of the state in
//       shared memory with 16-byte cp.async copies (fully coalesced rows, XOR-swizzled slots),
//       every thread multiplies its 2^K-amplitude groups by the matrix -- which sits in the
//       kernel-parameter constant bank, so the FMAs take it as an immediate constant operand
//       and no load instruction is spent on it -- and the tile is written back with 16-byte
//       coalesced stores.  Exactly one HBM read and one HBM write of the state.
//   pass_kernel<Real>
//       same staging, but a run of up to TCB200_MAX_PASS_OPS blocks is applied to the tile
//       before it is written back (matrices come from global memory via shared memory).
//   diag_kernel<Real>
//       streaming multiply by a 2^k-entry diagonal table.
//
// CUDA cores only: the path is HBM-bound (16 B of traffic per complex64 amplitude against
// 8*2^K flop), see DESIGN.md.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace tcb {

// ------------------------------------------------------------------------------------------------
// dense block, matrix in the parameter bank
// ------------------------------------------------------------------------------------------------
template <typename Real, int K>
struct DenseParams {
    typename CT<Real>::type* state;
    const typename CT<Real>::type* mats;  // BATCHED: [batch][4^K] device matrices
    TileGeom g;
    GroupMap gm;
    int tb;  // log2(blockDim.x)
    typename CT<Real>::type m[(1 << K) * (1 << K)];
};

template <typename Real, int K, bool BATCHED>
__global__ void __launch_bounds__(256) dense_kernel(const __grid_constant__ DenseParams<Real, K> p) {
    using C = typename CT<Real>::type;
    constexpr int D = 1 << K;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    C* tile = reinterpret_cast<C*>(smem_raw);
    __shared__ uint64_t rowoff[256];

    const int tid = threadIdx.x;
    const int nthr = blockDim.x;
    if (tid < (1 << p.g.h)) rowoff[tid] = row_offset(p.g, tid);
    C* vec = p.state + ((uint64_t)blockIdx.y << p.g.n);
    const uint64_t base = tile_base(p.g, blockIdx.x);

    C* bm = nullptr;
    if (BATCHED) {  // this batch element's matrix -> shared memory, behind the tile
        bm = tile + (1u << p.g.T);
        const C* src = p.mats + (uint64_t)blockIdx.y * (D * D);
        for (int i = tid; i < D * D; i += nthr) bm[i] = src[i];
    }
    __syncthreads();
    stage_in<C, true>(p.g, vec, base, tile, rowoff, tid, nthr);
    cp_async_wait_all();
    __syncthreads();

    // (the generic group loop is already at the HBM roofline for k <= 3 here: 40 registers,
    // 4 CTAs / SM; the unrolled fast path of the pass kernels would cost occupancy)
    if (BATCHED) {
        apply_block_on_tile<C, K>(tile, p.gm, tid, nthr, p.tb,
                                  [&](int i, int j) { return bm[i * D + j]; });
    } else {
        apply_block_on_tile<C, K>(tile, p.gm, tid, nthr, p.tb,
                                  [&](int i, int j) { return p.m[i * D + j]; });
    }
    __syncthreads();
    stage_out<C, true>(p.g, vec, base, tile, rowoff, tid, nthr);
}

template <typename Real, int K>
static int launch_dense(void* state, int nbits, const int* bits, const double* mat,
                        const void* mats_dev, int64_t batch, cudaStream_t st) {
    using C = typename CT<Real>::type;
    constexpr int D = 1 << K;
    constexpr int APU = CT<Real>::APU;
    // up to 16 KiB of parameters: keep the staging copy off the stack
    static thread_local DenseParams<Real, K>* tp = nullptr;
    if (!tp) tp = new DenseParams<Real, K>();
    DenseParams<Real, K>& q = *tp;
    q.state = static_cast<C*>(state);
    q.mats = static_cast<const C*>(mats_dev);
    const int tile_bits = dense_tile_bits(sizeof(Real) == 4 ? TCB200_C64 : TCB200_C128, K);
    int rc = make_geom(nbits, tile_bits, K, bits, &q.g);
    if (rc) return rc;
    rc = make_group_map(q.g, APU, K, bits, &q.gm);
    if (rc) return rc;
    q.tb = pick_threads(q.g.T, K, APU);
    if (!mats_dev) {
        for (int i = 0; i < D * D; ++i) {
            q.m[i].x = (Real)mat[2 * i];
            q.m[i].y = (Real)mat[2 * i + 1];
        }
    }
    const uint64_t ntiles = 1ull << (nbits - q.g.T);
    if (ntiles > 0x7fffffffull) return fail(TCB200_ERR_UNSUPPORTED, "state too large for one grid");
    if (batch < 1 || batch > 65535) return fail(TCB200_ERR_ARG, "batch=%lld out of range", (long long)batch);
    size_t smem = ((size_t)sizeof(C) << q.g.T);
    if (smem < 16) smem = 16;
    if (mats_dev) smem += sizeof(C) * D * D;
    dim3 grid((unsigned)ntiles, (unsigned)batch);
    dim3 block(1u << q.tb);
    if (mats_dev) {
        static bool attr = false;
        if (!attr) {
            TCB_CUDA(cudaFuncSetAttribute(dense_kernel<Real, K, true>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
            attr = true;
        }
        dense_kernel<Real, K, true><<<grid, block, smem, st>>>(q);
    } else {
        static bool attr = false;
        if (!attr) {
            TCB_CUDA(cudaFuncSetAttribute(dense_kernel<Real, K, false>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
            attr = true;
        }
        dense_kernel<Real, K, false><<<grid, block, smem, st>>>(q);
    }
    TCB_LAUNCH_CHECK("dense_kernel");
    return 0;
}

template <typename Real>
static int dispatch_dense(void* state, int nbits, int k, const int* bits, const double* mat,
                          const void* mats_dev, int64_t batch, cudaStream_t st) {
    switch (k) {
        case 1: return launch_dense<Real, 1>(state, nbits, bits, mat, mats_dev, batch, st);
        case 2: return launch_dense<Real, 2>(state, nbits, bits, mat, mats_dev, batch, st);
        case 3: return launch_dense<Real, 3>(state, nbits, bits, mat, mats_dev, batch, st);
        case 4: return launch_dense<Real, 4>(state, nbits, bits, mat, mats_dev, batch, st);
        case 5: return launch_dense<Real, 5>(state, nbits, bits, mat, mats_dev, batch, st);
    }
    return fail(TCB200_ERR_UNSUPPORTED, "dense block of %d bits unsupported (max %d)", k, TCB200_MAX_K);
}

static int check_common(const void* state, int nbits, int dtype, int k, const int* bits) {
    if (!state) return fail(TCB200_ERR_ARG, "state is NULL");
    if (dtype != TCB200_C64 && dtype != TCB200_C128) return fail(TCB200_ERR_ARG, "bad dtype %d", dtype);
    if (nbits < 1 || nbits > 40) return fail(TCB200_ERR_ARG, "nbits=%d out of range", nbits);
    if (k < 1) return fail(TCB200_ERR_ARG, "k=%d", k);
    if (!bits) return fail(TCB200_ERR_ARG, "bits is NULL");
    if (k > nbits) return fail(TCB200_ERR_ARG, "k=%d exceeds nbits=%d", k, nbits);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// multi-block pass
// ------------------------------------------------------------------------------------------------
struct PassParams {
    void* state;
    const void* mats;      // device blob, per op [batch_mats][4^k]
    int64_t mat_total;     // complex elements per batch element (sum 4^k)
    int batched_mats;      // 0: shared matrices, 1: per batch element
    TileGeom g;
    int tb;
    int nops;
    PassOp op[TCB200_MAX_PASS_OPS];
};

template <typename Real>
__global__ void __launch_bounds__(256) pass_kernel(const __grid_constant__ PassParams p) {
    using C = typename CT<Real>::type;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    C* tile = reinterpret_cast<C*>(smem_raw);
    C* bm = tile + (1u << p.g.T);
    __shared__ uint64_t rowoff[256];

    const int tid = threadIdx.x;
    const int nthr = blockDim.x;
    if (tid < (1 << p.g.h)) rowoff[tid] = row_offset(p.g, tid);
    C* vec = static_cast<C*>(p.state) + ((uint64_t)blockIdx.y << p.g.n);
    const uint64_t base = tile_base(p.g, blockIdx.x);
    {
        const C* src = static_cast<const C*>(p.mats);
        if (p.batched_mats) {
            // blob layout: op-major, each op holding [batch][4^k]
            for (int o = 0; o < p.nops; ++o) {
                const int sz = 1 << (2 * p.op[o].k);
                const C* s = src + (uint64_t)p.op[o].moff * gridDim.y + (uint64_t)blockIdx.y * sz;
                for (int i = tid; i < sz; i += nthr) bm[p.op[o].moff + i] = s[i];
            }
        } else {
            for (int i = tid; i < (int)p.mat_total; i += nthr) bm[i] = src[i];
        }
    }
    __syncthreads();
    stage_in<C, true>(p.g, vec, base, tile, rowoff, tid, nthr);
    cp_async_wait_all();
    __syncthreads();

    for (int o = 0; o < p.nops; ++o) {
        const PassOp& op = p.op[o];
        const C* m = bm + op.moff;
        switch (op.k) {
            case 1: {  // narrow blocks: matrix in registers for all the groups of the thread
                C r[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) r[i] = m[i];
                apply_block_dispatch<C,1>(tile, op.gm, tid, nthr, p.tb, [&](int i, int j) { return r[i * 2 + j]; });
                break;
            }
            case 2: {
                C r[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) r[i] = m[i];
                apply_block_dispatch<C,2>(tile, op.gm, tid, nthr, p.tb, [&](int i, int j) { return r[i * 4 + j]; });
                break;
            }
            case 3:
                apply_block_dispatch<C,3>(tile, op.gm, tid, nthr, p.tb, [&](int i, int j) { return m[i * 8 + j]; });
                break;
            default:
                apply_block_dispatch<C,4>(tile, op.gm, tid, nthr, p.tb, [&](int i, int j) { return m[i * 16 + j]; });
                break;
        }
        __syncthreads();
    }
    stage_out<C, true>(p.g, vec, base, tile, rowoff, tid, nthr);
}

template <typename Real>
static int launch_pass(void* state, int nbits, int nops, const int* ops_k, const int* ops_bits,
                       const void* mats, int64_t batch_mats, int n_hi, const int* tile_hi,
                       int64_t batch, cudaStream_t st) {
    using C = typename CT<Real>::type;
    constexpr int APU = CT<Real>::APU;
    static thread_local PassParams* tp = nullptr;
    if (!tp) tp = new PassParams();
    PassParams& q = *tp;
    q.state = state;
    q.mats = mats;
    q.batched_mats = batch_mats > 1 ? 1 : 0;
    if (batch_mats != 1 && batch_mats != batch)
        return fail(TCB200_ERR_ARG, "batch_mats must be 1 or batch");
    const int tile_bits = pass_tile_bits(sizeof(Real) == 4 ? TCB200_C64 : TCB200_C128);
    int rc = make_geom_hi(nbits, tile_bits, nbits <= tile_bits ? 0 : n_hi, tile_hi, &q.g);
    if (rc) return rc;
    q.nops = nops;
    int kmax = 0, kmin = 99;
    int64_t moff = 0;
    const int* b = ops_bits;
    for (int o = 0; o < nops; ++o) {
        const int k = ops_k[o];
        if (k < 1 || k > TCB200_MAX_PASS_K)
            return fail(TCB200_ERR_UNSUPPORTED, "block of %d bits inside a pass (max %d)", k, TCB200_MAX_PASS_K);
        for (int i = 0; i < k; ++i)
            if (b[i] < 0 || b[i] >= nbits) return fail(TCB200_ERR_ARG, "bit %d out of range", b[i]);
        q.op[o].k = k;
        q.op[o].moff = (int)moff;
        rc = make_group_map(q.g, APU, k, b, &q.op[o].gm);
        if (rc) return rc;
        moff += 1ll << (2 * k);
        b += k;
        if (k > kmax) kmax = k;
        if (k < kmin) kmin = k;
    }
    q.mat_total = moff;
    q.tb = pick_threads(q.g.T, kmin, APU);
    const uint64_t ntiles = 1ull << (nbits - q.g.T);
    if (ntiles > 0x7fffffffull) return fail(TCB200_ERR_UNSUPPORTED, "state too large for one grid");
    if (batch < 1 || batch > 65535) return fail(TCB200_ERR_ARG, "batch=%lld out of range", (long long)batch);
    size_t smem = ((size_t)sizeof(C) << q.g.T) + sizeof(C) * (size_t)moff;
    if (smem > 200 * 1024) return fail(TCB200_ERR_UNSUPPORTED, "pass needs %zu B of shared memory", smem);
    static bool attr = false;
    if (!attr) {
        TCB_CUDA(cudaFuncSetAttribute(pass_kernel<Real>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      200 * 1024));
        attr = true;
    }
    dim3 grid((unsigned)ntiles, (unsigned)batch);
    dim3 block(1u << q.tb);
    pass_kernel<Real><<<grid, block, smem, st>>>(q);
    TCB_LAUNCH_CHECK("pass_kernel");
    return 0;
}

// ------------------------------------------------------------------------------------------------
// multi-block pass, shared matrices in the parameter constant bank
// ------------------------------------------------------------------------------------------------
constexpr int CPASS_MAT_BYTES = 12 * 1024;

template <typename Real>
struct CPassParams {
    typename CT<Real>::type* state;
    TileGeom g;
    int tb;
    int nops;
    int mat_total;  // complex elements used in m[]
    PassOp op[TCB200_MAX_PASS_OPS];
    typename CT<Real>::type m[CPASS_MAT_BYTES / sizeof(typename CT<Real>::type)];
};

// the blocks of the pass on one staged tile (a __syncthreads() after every block)
template <typename Real>
__device__ __forceinline__ void cpass_ops(const CPassParams<Real>& p, typename CT<Real>::type* tile,
                                          const typename CT<Real>::type* bm, int tid, int nthr) {
    using C = typename CT<Real>::type;
    for (int o = 0; o < p.nops; ++o) {
        const PassOp& op = p.op[o];
        const C* m = bm + op.moff;
        switch (op.k) {
            case 1: {
                C r[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) r[i] = m[i];
                apply_block_dispatch<C,1>(tile, op.gm, tid, nthr, p.tb, [&](int i, int j) { return r[i * 2 + j]; });
                break;
            }
            case 2: {
                C r[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) r[i] = m[i];
                apply_block_dispatch<C,2>(tile, op.gm, tid, nthr, p.tb, [&](int i, int j) { return r[i * 4 + j]; });
                break;
            }
            case 3:
                apply_block_dispatch<C,3>(tile, op.gm, tid, nthr, p.tb, [&](int i, int j) { return m[i * 8 + j]; });
                break;
            default:
                apply_block_dispatch<C,4>(tile, op.gm, tid, nthr, p.tb, [&](int i, int j) { return m[i * 16 + j]; });
                break;
        }
        __syncthreads();
    }
}

template <typename Real>
__global__ void __launch_bounds__(256) cpass_kernel(const __grid_constant__ CPassParams<Real> p) {
    using C = typename CT<Real>::type;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    C* tile = reinterpret_cast<C*>(smem_raw);
    __shared__ uint64_t rowoff[256];

    const int tid = threadIdx.x;
    const int nthr = blockDim.x;
    if (tid < (1 << p.g.h)) rowoff[tid] = row_offset(p.g, tid);
    C* vec = p.state + ((uint64_t)blockIdx.y << p.g.n);
    const uint64_t base = tile_base(p.g, blockIdx.x);
    // The constant bank is only the transport of the matrices: an index that depends on the
    // op counter would make every FMA operand a register-indexed LDC.  Copy them once per CTA
    // behind the tile; 1- and 2-bit blocks then keep their matrix in registers for all the
    // groups of a thread, wider blocks read it with broadcast LDS.
    C* bm = tile + (1u << p.g.T);
    for (int i = tid; i < p.mat_total; i += nthr) bm[i] = p.m[i];
    __syncthreads();
    stage_in<C, true>(p.g, vec, base, tile, rowoff, tid, nthr);
    cp_async_wait_all();
    __syncthreads();

    cpass_ops<Real>(p, tile, bm, tid, nthr);
    stage_out<C, true>(p.g, vec, base, tile, rowoff, tid, nthr);
}

template <typename Real>
static int launch_cpass(void* state, int nbits, int nops, const int* ops_k, const int* ops_bits,
                        const double* mats, int n_hi, const int* tile_hi, int64_t batch, cudaStream_t st) {
    using C = typename CT<Real>::type;
    constexpr int APU = CT<Real>::APU;
    constexpr int MAXM = CPASS_MAT_BYTES / (int)sizeof(C);
    static thread_local CPassParams<Real>* tp = nullptr;
    if (!tp) tp = new CPassParams<Real>();
    CPassParams<Real>& q = *tp;
    q.state = static_cast<C*>(state);
    const int tile_bits = pass_tile_bits(sizeof(Real) == 4 ? TCB200_C64 : TCB200_C128);
    int rc = make_geom_hi(nbits, tile_bits, nbits <= tile_bits ? 0 : n_hi, tile_hi, &q.g);
    if (rc) return rc;
    q.nops = nops;
    int kmin = 99;
    int moff = 0;
    const int* b = ops_bits;
    const double* mp = mats;
    for (int o = 0; o < nops; ++o) {
        const int k = ops_k[o];
        if (k < 1 || k > TCB200_MAX_PASS_K)
            return fail(TCB200_ERR_UNSUPPORTED, "block of %d bits inside a pass (max %d)", k, TCB200_MAX_PASS_K);
        for (int i = 0; i < k; ++i)
            if (b[i] < 0 || b[i] >= nbits) return fail(TCB200_ERR_ARG, "bit %d out of range", b[i]);
        const int sz = 1 << (2 * k);
        if (moff + sz > MAXM) return fail(TCB200_ERR_UNSUPPORTED, "pass matrices exceed %d bytes", CPASS_MAT_BYTES);
        q.op[o].k = k;
        q.op[o].moff = moff;
        rc = make_group_map(q.g, APU, k, b, &q.op[o].gm);
        if (rc) return rc;
        for (int i = 0; i < sz; ++i) {
            q.m[moff + i].x = (Real)mp[2 * i];
            q.m[moff + i].y = (Real)mp[2 * i + 1];
        }
        moff += sz;
        mp += 2 * sz;
        b += k;
        if (k < kmin) kmin = k;
    }
    q.mat_total = moff;
    q.tb = pick_threads(q.g.T, kmin, APU);
    const uint64_t ntiles = 1ull << (nbits - q.g.T);
    if (ntiles > 0x7fffffffull) return fail(TCB200_ERR_UNSUPPORTED, "state too large for one grid");
    if (batch < 1 || batch > 65535) return fail(TCB200_ERR_ARG, "batch=%lld out of range", (long long)batch);
    size_t smem = ((size_t)sizeof(C) << q.g.T) + sizeof(C) * (size_t)moff;
    if (smem < 16) smem = 16;
    static bool attr = false;
    if (!attr) {
        TCB_CUDA(cudaFuncSetAttribute(cpass_kernel<Real>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr = true;
    }
    dim3 grid((unsigned)ntiles, (unsigned)batch);
    dim3 block(1u << q.tb);
    cpass_kernel<Real><<<grid, block, smem, st>>>(q);
    TCB_LAUNCH_CHECK("cpass_kernel");
    return 0;
}

// ------------------------------------------------------------------------------------------------
// register-tile pass: per shared-memory round trip a *sequence* of 1-/2-bit gates inside a
// <= 4-bit register tile (common.cuh: rtile_run)
// ------------------------------------------------------------------------------------------------
constexpr int RPASS_MAX_SUB = 96;

template <typename Real>
struct RPassParams {
    typename CT<Real>::type* state;
    TileGeom g;
    int tb;
    int nrt;
    int mat_total;
    RTile rt[TCB200_MAX_PASS_OPS];
    RSub sub[RPASS_MAX_SUB];
    typename CT<Real>::type m[CPASS_MAT_BYTES / sizeof(typename CT<Real>::type)];
};

template <typename Real>
__global__ void __launch_bounds__(256, 3) rpass_kernel(const __grid_constant__ RPassParams<Real> p) {
    using C = typename CT<Real>::type;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    C* tile = reinterpret_cast<C*>(smem_raw);
    __shared__ uint64_t rowoff[256];
    __shared__ RSub subs[RPASS_MAX_SUB];

    const int tid = threadIdx.x;
    const int nthr = blockDim.x;
    if (tid < (1 << p.g.h)) rowoff[tid] = row_offset(p.g, tid);
    C* vec = p.state + ((uint64_t)blockIdx.y << p.g.n);
    const uint64_t base = tile_base(p.g, blockIdx.x);
    C* bm = tile + (1u << p.g.T);
    for (int i = tid; i < p.mat_total; i += nthr) bm[i] = p.m[i];
    for (int i = tid; i < RPASS_MAX_SUB; i += nthr) subs[i] = p.sub[i];
    __syncthreads();
    stage_in<C, true>(p.g, vec, base, tile, rowoff, tid, nthr);
    cp_async_wait_all();
    __syncthreads();
    for (int o = 0; o < p.nrt; ++o) {
        rtile_run<C>(tile, p.rt[o], subs, bm, tid, nthr, p.tb);
        __syncthreads();
    }
    stage_out<C, true>(p.g, vec, base, tile, rowoff, tid, nthr);
}

// host side: the parameter block of one register-tile pass
template <typename Real>
static int fill_rpass(RPassParams<Real>& q, void* state, int nbits, int nrt, const int* rt_k, const int* rt_bits,
                      const int* rt_nsub, const int* sub_k, const int* sub_bits, const double* sub_mats, int n_hi,
                      const int* tile_hi) {
    using C = typename CT<Real>::type;
    constexpr int APU = CT<Real>::APU;
    constexpr int MAXM = CPASS_MAT_BYTES / (int)sizeof(C);
    q.state = static_cast<C*>(state);
    const int tile_bits = pass_tile_bits(sizeof(Real) == 4 ? TCB200_C64 : TCB200_C128);
    int rc = make_geom_hi(nbits, tile_bits, nbits <= tile_bits ? 0 : n_hi, tile_hi, &q.g);
    if (rc) return rc;
    q.nrt = nrt;
    int moff = 0, nsub_tot = 0, kmin = 99;
    const int* rb = rt_bits;
    const int* sb = sub_bits;
    const double* mp = sub_mats;
    const int* sk = sub_k;
    for (int o = 0; o < nrt; ++o) {
        const int kt = rt_k[o];
        if (kt < 1 || kt > (sizeof(Real) == 4 ? 4 : 3))
            return fail(TCB200_ERR_UNSUPPORTED, "register tile of %d bits (max 4 for complex64, 3 for complex128)", kt);
        for (int i = 0; i < kt; ++i)
            if (rb[i] < 0 || rb[i] >= nbits) return fail(TCB200_ERR_ARG, "bit %d out of range", rb[i]);
        q.rt[o].kt = kt;
        q.rt[o].nsub = rt_nsub[o];
        q.rt[o].sub0 = nsub_tot;
        rc = make_group_map(q.g, APU, kt, rb, &q.rt[o].gm);
        if (rc) return rc;
        for (int s = 0; s < rt_nsub[o]; ++s) {
            if (nsub_tot >= RPASS_MAX_SUB) return fail(TCB200_ERR_UNSUPPORTED, "more than %d gates in one pass", RPASS_MAX_SUB);
            const int k = *sk++;
            if (k != 1 && k != 2) return fail(TCB200_ERR_UNSUPPORTED, "gate of %d bits inside a register tile", k);
            int pos[2] = {0, 0};
            for (int i = 0; i < k; ++i) {
                int f = -1;
                for (int j = 0; j < kt; ++j)
                    if (rb[j] == sb[i]) f = j;
                if (f < 0) return fail(TCB200_ERR_ARG, "gate bit %d is not in its register tile", sb[i]);
                pos[i] = f;
            }
            if (k == 2 && pos[1] <= pos[0]) return fail(TCB200_ERR_ARG, "gate bits must be ascending");
            const int sz = 1 << (2 * k);
            if (moff + sz > MAXM) return fail(TCB200_ERR_UNSUPPORTED, "pass matrices exceed %d bytes", CPASS_MAT_BYTES);
            RSub& r = q.sub[nsub_tot++];
            r.k = k;
            r.p0 = pos[0];
            r.p1 = pos[1];
            r.moff = moff;
            for (int i = 0; i < sz; ++i) {
                q.m[moff + i].x = (Real)mp[2 * i];
                q.m[moff + i].y = (Real)mp[2 * i + 1];
            }
            moff += sz;
            mp += 2 * sz;
            sb += k;
        }
        rb += kt;
        if (kt < kmin) kmin = kt;
    }
    q.mat_total = moff;
    q.tb = pick_threads(q.g.T, kmin, APU);
    return 0;
}

template <typename Real>
static int launch_rpass(void* state, int nbits, int nrt, const int* rt_k, const int* rt_bits, const int* rt_nsub,
                        const int* sub_k, const int* sub_bits, const double* sub_mats, int n_hi, const int* tile_hi,
                        int64_t batch, cudaStream_t st) {
    using C = typename CT<Real>::type;
    static thread_local RPassParams<Real>* tp = nullptr;
    if (!tp) tp = new RPassParams<Real>();
    RPassParams<Real>& q = *tp;
    int rc = fill_rpass<Real>(q, state, nbits, nrt, rt_k, rt_bits, rt_nsub, sub_k, sub_bits, sub_mats, n_hi, tile_hi);
    if (rc) return rc;
    const uint64_t ntiles = 1ull << (nbits - q.g.T);
    if (ntiles > 0x7fffffffull) return fail(TCB200_ERR_UNSUPPORTED, "state too large for one grid");
    if (batch < 1 || batch > 65535) return fail(TCB200_ERR_ARG, "batch=%lld out of range", (long long)batch);
    size_t smem = ((size_t)sizeof(C) << q.g.T) + sizeof(C) * (size_t)q.mat_total;
    if (smem < 16) smem = 16;
    static bool attr = false;
    if (!attr) {
        TCB_CUDA(cudaFuncSetAttribute(rpass_kernel<Real>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr = true;
    }
    dim3 grid((unsigned)ntiles, (unsigned)batch);
    dim3 block(1u << q.tb);
    rpass_kernel<Real><<<grid, block, smem, st>>>(q);
    TCB_LAUNCH_CHECK("rpass_kernel");
    return 0;
}

#ifdef TCB200_EMU
// tests/emu only (never compiled into the product library): run the register-tile pass on the
// CPU with the same parameter block and the same __host__ __device__ bodies as rpass_kernel.
template <typename Real>
static int emu_rpass(void* state, int nbits, int nrt, const int* rt_k, const int* rt_bits, const int* rt_nsub,
                     const int* sub_k, const int* sub_bits, const double* sub_mats, int n_hi, const int* tile_hi) {
    using C = typename CT<Real>::type;
    RPassParams<Real>* q = new RPassParams<Real>();
    int rc = fill_rpass<Real>(*q, state, nbits, nrt, rt_k, rt_bits, rt_nsub, sub_k, sub_bits, sub_mats, n_hi, tile_hi);
    if (rc) { delete q; return rc; }
    const int nthr = 1 << q->tb;
    const size_t elems = ((size_t)1 << q->g.T) + q->mat_total + 16;
    C* tile = static_cast<C*>(aligned_alloc(128, ((elems * sizeof(C) + 127) / 128) * 128));
    C* bm = tile + ((size_t)1 << q->g.T);
    for (int i = 0; i < q->mat_total; ++i) bm[i] = q->m[i];
    uint64_t rowoff[256];
    for (int r = 0; r < (1 << q->g.h); ++r) rowoff[r] = row_offset(q->g, r);
    C* vec = static_cast<C*>(state);
    const uint64_t ntiles = 1ull << (nbits - q->g.T);
    for (uint64_t t = 0; t < ntiles; ++t) {
        const uint64_t base = tile_base(q->g, t);
        for (int tid = 0; tid < nthr; ++tid) stage_in<C, true>(q->g, vec, base, tile, rowoff, tid, nthr);
        for (int o = 0; o < q->nrt; ++o)
            for (int tid = 0; tid < nthr; ++tid) rtile_run<C>(tile, q->rt[o], q->sub, bm, tid, nthr, q->tb);
        for (int tid = 0; tid < nthr; ++tid) stage_out<C, true>(q->g, vec, base, tile, rowoff, tid, nthr);
    }
    free(tile);
    delete q;
    return 0;
}

extern "C" __attribute__((visibility("default"))) int emu_apply_rpass(void* state, int nbits, int dtype, int nrt, const int* rt_k,
                                                                       const int* rt_bits, const int* rt_nsub, const int* sub_k,
                                                                       const int* sub_bits, const double* sub_mats, int n_hi,
                                                                       const int* tile_hi) {
    if (dtype == TCB200_C64) return emu_rpass<float>(state, nbits, nrt, rt_k, rt_bits, rt_nsub, sub_k, sub_bits, sub_mats, n_hi, tile_hi);
    return emu_rpass<double>(state, nbits, nrt, rt_k, rt_bits, rt_nsub, sub_k, sub_bits, sub_mats, n_hi, tile_hi);
}
#endif

// ------------------------------------------------------------------------------------------------
// diagonal block
// ------------------------------------------------------------------------------------------------
struct DiagParams {
    void* state;
    const void* table;  // device [batch or 1][2^k]
    int batched;
    int n;
    int k;
    int bits[TCB200_MAX_DIAG_K];
    uint64_t nunits;  // 16-byte units per state vector
};

template <typename Real>
__global__ void __launch_bounds__(256) diag_kernel(const __grid_constant__ DiagParams p) {
    using C = typename CT<Real>::type;
    constexpr int APU = CT<Real>::APU;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    C* tab = reinterpret_cast<C*>(smem_raw);
    const int tsz = 1 << p.k;
    const C* src = static_cast<const C*>(p.table) + (p.batched ? (uint64_t)blockIdx.y * tsz : 0);
    for (int i = threadIdx.x; i < tsz; i += blockDim.x) tab[i] = src[i];
    __syncthreads();
    C* vec = static_cast<C*>(p.state) + ((uint64_t)blockIdx.y << p.n);
    Unit16* v16 = reinterpret_cast<Unit16*>(vec);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; u < p.nunits; u += stride) {
        Unit16 q = v16[u];
        C* qc = reinterpret_cast<C*>(&q);
        const uint64_t r0 = u * APU;
#pragma unroll
        for (int a = 0; a < APU; ++a) {
            const uint64_t r = r0 + a;
            uint32_t idx = 0;
            for (int j = 0; j < p.k; ++j) idx |= (uint32_t)((r >> p.bits[j]) & 1ull) << j;
            qc[a] = cmul(tab[idx], qc[a]);
        }
        v16[u] = q;
    }
}

}  // namespace tcb

using namespace tcb;

extern "C" {

int tcb200_apply_dense(void* state, int nbits, int dtype, int k, const int* bits,
                       const double* mat, int64_t batch, void* stream) {
    int rc = check_common(state, nbits, dtype, k, bits);
    if (rc) return rc;
    if (!mat) return fail(TCB200_ERR_ARG, "mat is NULL");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == TCB200_C64) return dispatch_dense<float>(state, nbits, k, bits, mat, nullptr, batch, st);
    return dispatch_dense<double>(state, nbits, k, bits, mat, nullptr, batch, st);
}

int tcb200_apply_dense_batched(void* state, int nbits, int dtype, int k, const int* bits,
                               const void* mats_dev, int64_t batch, void* stream) {
    int rc = check_common(state, nbits, dtype, k, bits);
    if (rc) return rc;
    if (!mats_dev) return fail(TCB200_ERR_ARG, "mats_dev is NULL");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == TCB200_C64) return dispatch_dense<float>(state, nbits, k, bits, nullptr, mats_dev, batch, st);
    return dispatch_dense<double>(state, nbits, k, bits, nullptr, mats_dev, batch, st);
}

int tcb200_pass_tile_bits(int dtype) { return pass_tile_bits(dtype); }

int tcb200_apply_pass(void* state, int nbits, int dtype, int nops, const int* ops_k,
                      const int* ops_bits, const void* ops_mats_dev, int64_t batch_mats,
                      int n_hi, const int* tile_hi, int64_t batch, void* stream) {
    if (!state || !ops_k || !ops_bits || !ops_mats_dev) return fail(TCB200_ERR_ARG, "NULL argument");
    if (dtype != TCB200_C64 && dtype != TCB200_C128) return fail(TCB200_ERR_ARG, "bad dtype %d", dtype);
    if (nops < 1 || nops > TCB200_MAX_PASS_OPS) return fail(TCB200_ERR_ARG, "nops=%d out of range", nops);
    if (nbits < 1 || nbits > 40) return fail(TCB200_ERR_ARG, "nbits=%d out of range", nbits);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == TCB200_C64)
        return launch_pass<float>(state, nbits, nops, ops_k, ops_bits, ops_mats_dev, batch_mats, n_hi, tile_hi, batch, st);
    return launch_pass<double>(state, nbits, nops, ops_k, ops_bits, ops_mats_dev, batch_mats, n_hi, tile_hi, batch, st);
}

int tcb200_apply_pass_host(void* state, int nbits, int dtype, int nops, const int* ops_k,
                           const int* ops_bits, const double* ops_mats, int n_hi,
                           const int* tile_hi, int64_t batch, void* stream) {
    if (!state || !ops_k || !ops_bits || !ops_mats) return fail(TCB200_ERR_ARG, "NULL argument");
    if (dtype != TCB200_C64 && dtype != TCB200_C128) return fail(TCB200_ERR_ARG, "bad dtype %d", dtype);
    if (nops < 1 || nops > TCB200_MAX_PASS_OPS) return fail(TCB200_ERR_ARG, "nops=%d out of range", nops);
    if (nbits < 1 || nbits > 40) return fail(TCB200_ERR_ARG, "nbits=%d out of range", nbits);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int tr = launch_tpass(dtype, state, nbits, nops, ops_k, ops_bits, ops_mats, n_hi, tile_hi, batch, st);
    if (tr <= 0) return tr;  // launched (0) or failed (< 0); > 0: not eligible for the TMA pipeline
    if (dtype == TCB200_C64)
        return launch_cpass<float>(state, nbits, nops, ops_k, ops_bits, ops_mats, n_hi, tile_hi, batch, st);
    return launch_cpass<double>(state, nbits, nops, ops_k, ops_bits, ops_mats, n_hi, tile_hi, batch, st);
}

int tcb200_apply_rpass_host(void* state, int nbits, int dtype, int nrt, const int* rt_k,
                            const int* rt_bits, const int* rt_nsub, const int* sub_k,
                            const int* sub_bits, const double* sub_mats, int n_hi,
                            const int* tile_hi, int64_t batch, void* stream) {
    if (!state || !rt_k || !rt_bits || !rt_nsub || !sub_k || !sub_bits || !sub_mats) return fail(TCB200_ERR_ARG, "NULL argument");
    if (dtype != TCB200_C64 && dtype != TCB200_C128) return fail(TCB200_ERR_ARG, "bad dtype %d", dtype);
    if (nrt < 1 || nrt > TCB200_MAX_PASS_OPS) return fail(TCB200_ERR_ARG, "nrt=%d out of range", nrt);
    if (nbits < 1 || nbits > 40) return fail(TCB200_ERR_ARG, "nbits=%d out of range", nbits);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == TCB200_C64) {
        const int tr = launch_trpass(state, nbits, nrt, rt_k, rt_bits, rt_nsub, sub_k, sub_bits, sub_mats, n_hi, tile_hi, batch, st);
        if (tr <= 0) return tr;  // launched or failed; > 0: not eligible for the TMA pipeline
    }
    if (dtype == TCB200_C64)
        return launch_rpass<float>(state, nbits, nrt, rt_k, rt_bits, rt_nsub, sub_k, sub_bits, sub_mats, n_hi, tile_hi, batch, st);
    return launch_rpass<double>(state, nbits, nrt, rt_k, rt_bits, rt_nsub, sub_k, sub_bits, sub_mats, n_hi, tile_hi, batch, st);
}

int tcb200_apply_diag(void* state, int nbits, int dtype, int k, const int* bits,
                      const double* diag, const void* diag_dev, int64_t batch, void* stream) {
    int rc = check_common(state, nbits, dtype, k, bits);
    if (rc) return rc;
    if (k > TCB200_MAX_DIAG_K) return fail(TCB200_ERR_UNSUPPORTED, "diagonal block of %d bits", k);
    if (!diag && !diag_dev) return fail(TCB200_ERR_ARG, "diag is NULL");
    if (batch < 1 || batch > 65535) return fail(TCB200_ERR_ARG, "batch=%lld out of range", (long long)batch);
    for (int i = 0; i < k; ++i)
        if (bits[i] < 0 || bits[i] >= nbits || (i > 0 && bits[i] <= bits[i - 1]))
            return fail(TCB200_ERR_ARG, "bad bit list");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DiagParams p;
    memset(&p, 0, sizeof(p));
    p.state = state;
    p.n = nbits;
    p.k = k;
    for (int i = 0; i < k; ++i) p.bits[i] = bits[i];
    const size_t esz = dtype == TCB200_C64 ? 8 : 16;
    p.nunits = ((uint64_t)esz << nbits) / 16;
    if (p.nunits == 0) return fail(TCB200_ERR_UNSUPPORTED, "state smaller than 16 bytes");
    void* staged = nullptr;
    if (diag_dev) {
        p.table = diag_dev;
        p.batched = 1;
    } else {
        // small table: stage through a stream-ordered allocation
        const size_t bytes = esz << k;
        TCB_CUDA(cudaMallocAsync(&staged, bytes, st));
        static thread_local unsigned char* hbuf = nullptr;
        if (!hbuf) TCB_CUDA(cudaMallocHost(&hbuf, 16u << TCB200_MAX_DIAG_K));
        for (int i = 0; i < (1 << k); ++i) {
            if (dtype == TCB200_C64) {
                reinterpret_cast<float*>(hbuf)[2 * i] = (float)diag[2 * i];
                reinterpret_cast<float*>(hbuf)[2 * i + 1] = (float)diag[2 * i + 1];
            } else {
                reinterpret_cast<double*>(hbuf)[2 * i] = diag[2 * i];
                reinterpret_cast<double*>(hbuf)[2 * i + 1] = diag[2 * i + 1];
            }
        }
        TCB_CUDA(cudaMemcpyAsync(staged, hbuf, bytes, cudaMemcpyHostToDevice, st));
        TCB_CUDA(cudaStreamSynchronize(st));  // hbuf is reused by the next call
        p.table = staged;
        p.batched = 0;
    }
    uint64_t want = (p.nunits + 255) / 256;
    unsigned gx = (unsigned)(want < 148ull * 16 ? (want ? want : 1) : 148ull * 16);
    dim3 grid(gx, (unsigned)batch);
    const size_t smem = esz << k;
    if (dtype == TCB200_C64) {
        static bool attr = false;
        if (!attr) { TCB_CUDA(cudaFuncSetAttribute(diag_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)); attr = true; }
        diag_kernel<float><<<grid, 256, smem, st>>>(p);
    } else {
        static bool attr = false;
        if (!attr) { TCB_CUDA(cudaFuncSetAttribute(diag_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)); attr = true; }
        diag_kernel<double><<<grid, 256, smem, st>>>(p);
    }
    TCB_LAUNCH_CHECK("diag_kernel");
    if (staged) TCB_CUDA(cudaFreeAsync(staged, st));
    return 0;
}

}  // extern "C"
