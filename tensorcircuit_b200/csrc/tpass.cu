// tpass_kernel: the multi-block pass as a persistent, warp-specialised TMA pipeline.
//
// cpass_kernel (apply.cu) stages a tile with LDGSTS, runs the blocks of the pass on it and
// writes it back -- three phases per CTA that only overlap statistically between the CTAs of an
// SM; measured at n = 34 a pass costs about (HBM time + FP32 time), not their maximum
// (DESIGN.md 4).  Here one CTA per SM stays resident and walks the tiles t = blockIdx.x + i *
// gridDim.x through a ring of three 64 KiB shared-memory buffers:
//
//   producer warp (one lane)                 16 compute warps
//   ---------------------------------        -------------------------------------------
//   TMA load  tile i+2  -> buf[(i+2)%3]      wait full[i%3]
//   wait done[i%3]                           blocks of the pass on buf[i%3] (bar.sync 1 between)
//   TMA store buf[i%3]  -> tile i            fence.proxy.async; arrive done[i%3]
//
// so the load of tile i+2 and the write-back of tile i-1 are in flight while tile i is being
// computed, with no load/store instruction issued by the compute warps.  Tiles travel as
// cp.async.bulk.tensor boxes (rank-5 tensor map over 128-byte lines, the gathered bits of the tile
// are stride dimensions -- common.cuh TmaPlan) and sit in shared memory in the hardware
// SWIZZLE_128B layout; the group maps are built for that layout (make_group_map, SWZ_HW128).
#include <cuda.h>  // CUtensorMap types; the encoder is looked up through cudart at run time
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

#ifdef TCB200_EMU
#include <algorithm>
#include <vector>
#endif

namespace tcb {

constexpr int TPASS_TILE_BYTES_LOG2 = PASS_TILE_BYTES_LOG2;   // 64 KiB: one ring per SM; 32 KiB: two
constexpr int TPASS_TB = TPASS_TILE_BYTES_LOG2 - 3 - 4;      // one 16-amplitude complex64 group per thread
constexpr int TPASS_NCOMPUTE = 1 << TPASS_TB;
constexpr int TPASS_CTAS_PER_SM = TPASS_TILE_BYTES_LOG2 >= 16 ? 1 : 2;
constexpr int TPASS_STAGES = 3;
constexpr int TPASS_MAT_BYTES = 12 * 1024;

template <typename Real>
struct TPassParams {
    TileGeom g;
    TmaPlan tp;
    int nops;
    int mat_total;      // complex elements used in m[]
    int vec_tiles_log2; // log2(tiles per state vector)
    uint64_t ntiles;    // over all vectors of the batch
    PassOp op[TCB200_MAX_PASS_OPS];
    typename CT<Real>::type m[TPASS_MAT_BYTES / sizeof(typename CT<Real>::type)];
};

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    const uint32_t a = smem_u32(b);
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(0), "r"(c1), "r"(0), "r"(0), "r"(0)
        : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, const void* src, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];\n"
        ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(0), "r"(c1), "r"(0), "r"(0), "r"(0)
        : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, %0;\n" ::"n"(TPASS_NCOMPUTE) : "memory"); }

// ---- producer: every bulk copy of the CTA is issued by this one lane ----------------------------
// Tile i of the CTA (global tile first + i * stride) lives in ring buffer i % 3.
__device__ __forceinline__ void tpass_producer(const CUtensorMap* tmap, const TileGeom& g, const TmaPlan& tp,
                                               int vec_tiles_log2, uint64_t nmine, uint64_t first, uint64_t stride,
                                               unsigned char* ring, uint64_t* full, uint64_t* done, int amp_bytes) {
    constexpr uint32_t TILE_BYTES = 1u << TPASS_TILE_BYTES_LOG2;
    const uint64_t vmask = (1ull << vec_tiles_log2) - 1ull;
    const uint32_t nbox = 1u << tp.nextra;
    const size_t box_bytes = ((size_t)amp_bytes) << tp.box_amps_log2;
    auto tile_amp = [&](uint64_t i) -> uint64_t {  // amplitude index of the tile's element 0
        const uint64_t t = first + i * stride;
        return ((t >> vec_tiles_log2) << g.n) + tile_base(g, t & vmask);
    };
    auto load = [&](uint64_t i) {
        const int s = (int)(i % TPASS_STAGES);
        unsigned char* buf = ring + s * TILE_BYTES;
        const uint64_t base = tile_amp(i);
        mbar_arrive_expect_tx(&full[s], TILE_BYTES);
        for (uint32_t c = 0; c < nbox; ++c)
            tma_load_5d(buf + c * box_bytes, tmap, &full[s], (int)((base + tma_box_offset(tp, c)) >> tp.line_bits));
    };
    if (nmine > 0) load(0);
    if (nmine > 1) load(1);
    for (uint64_t i = 0; i < nmine; ++i) {
        const int s = (int)(i % TPASS_STAGES);
        if (i + 2 < nmine) {
            // buf[(i+2)%3] was written back as tile i-1: its bulk store must have read it
            if (i >= 1) bulk_wait_read0();
            load(i + 2);
        }
        mbar_wait(&done[s], (uint32_t)((i / TPASS_STAGES) & 1));
        const unsigned char* buf = ring + s * TILE_BYTES;
        const uint64_t base = tile_amp(i);
        for (uint32_t c = 0; c < nbox; ++c)
            tma_store_5d(tmap, buf + c * box_bytes, (int)((base + tma_box_offset(tp, c)) >> tp.line_bits));
        bulk_commit();
    }
    bulk_wait0();
}

__device__ __forceinline__ void tpass_init_barriers(uint64_t* full, uint64_t* done) {
    for (int s = 0; s < TPASS_STAGES; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&done[s], TPASS_NCOMPUTE);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    fence_proxy_async();
}

template <typename Real>
__global__ void __launch_bounds__(TPASS_NCOMPUTE + 32, TPASS_CTAS_PER_SM)
tpass_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TPassParams<Real> p) {
    using C = typename CT<Real>::type;
    constexpr int T = TPASS_TILE_BYTES_LOG2 - (sizeof(Real) == 4 ? 3 : 4);
    constexpr uint32_t TILE_BYTES = 1u << TPASS_TILE_BYTES_LOG2;
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full[TPASS_STAGES];
    __shared__ __align__(8) uint64_t done[TPASS_STAGES];

    // SWIZZLE_128B atoms are 1024 bytes: align the ring by hand (the launch reserves the slack)
    unsigned char* ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    C* bm = reinterpret_cast<C*>(ring + TPASS_STAGES * TILE_BYTES);

    const int tid = threadIdx.x;
    if (tid == 0) tpass_init_barriers(full, done);
    for (int i = tid; i < p.mat_total; i += blockDim.x) bm[i] = p.m[i];
    __syncthreads();

    const uint64_t first = blockIdx.x, stride = gridDim.x;
    const uint64_t nmine = p.ntiles > first ? (p.ntiles - first + stride - 1) / stride : 0;

    if (tid >= TPASS_NCOMPUTE) {
        if (tid == TPASS_NCOMPUTE)
            tpass_producer(&tmap, p.g, p.tp, p.vec_tiles_log2, nmine, first, stride, ring, full, done, (int)sizeof(C));
        return;
    }

    // ---- compute warps ------------------------------------------------------------------------------
    for (uint64_t i = 0; i < nmine; ++i) {
        const int s = (int)(i % TPASS_STAGES);
        C* tile = reinterpret_cast<C*>(ring + s * TILE_BYTES);
        mbar_wait(&full[s], (uint32_t)((i / TPASS_STAGES) & 1));
        for (int o = 0; o < p.nops; ++o) {
            const PassOp& op = p.op[o];
            const C* m = bm + op.moff;
            switch (op.k) {
                case 1: {
                    C r[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) r[q] = m[q];
                    apply_block_tb<C, 1, T, TPASS_TB>(tile, op.gm, tid, [&](int a, int b) { return r[a * 2 + b]; });
                    break;
                }
                case 2: {
                    C r[16];
#pragma unroll
                    for (int q = 0; q < 16; ++q) r[q] = m[q];
                    apply_block_tb<C, 2, T, TPASS_TB>(tile, op.gm, tid, [&](int a, int b) { return r[a * 4 + b]; });
                    break;
                }
                case 3:
                    apply_block_tb<C, 3, T, TPASS_TB>(tile, op.gm, tid, [&](int a, int b) { return m[a * 8 + b]; });
                    break;
                default:
                    apply_block_tb<C, 4, T, TPASS_TB>(tile, op.gm, tid, [&](int a, int b) { return m[a * 16 + b]; });
                    break;
            }
            if (o + 1 < p.nops) bar_compute();
        }
        // generic-proxy writes of this thread -> visible to the bulk store, then hand the buffer over
        fence_proxy_async();
        mbar_arrive(&done[s]);
    }
}

// ---- complex64 passes of 1-/2-bit gates: 16-amplitude register tiles ---------------------------------
// Same pipeline; the compute warps run trtile_thread (common.cuh): per op one thread owns one
// group of 16 amplitudes for ALL the gates of the op, so the shared-memory traffic and the
// address arithmetic are paid once per op instead of once per gate.  The swizzled byte offset of
// every thread's group is tile-invariant: it is tabulated once per CTA (16-bit entries) when the
// kernel starts.  Ops of one segment (host plan below) are separated by __syncwarp() only.
constexpr int TRPASS_MAX_SUB = 48;
constexpr int TRPASS_MAT_ELEMS = TRPASS_MAX_SUB * 16;  // 6 KiB: one 16-element slot per gate

struct TRPassParams {
    TileGeom g;
    TmaPlan tp;
    int nops;
    int nsub;
    int mat_total;
    int vec_tiles_log2;
    uint64_t ntiles;
    TROp op[TCB200_MAX_PASS_OPS];
    uint32_t lane8[TCB200_MAX_PASS_OPS][TPASS_TB];  // byte offset contributed by thread-index bit i
    uint32_t subcode[TRPASS_MAX_SUB];               // gate case | matrix offset << 8 (common.cuh)
    float2 m[TRPASS_MAT_ELEMS];
};

__global__ void __launch_bounds__(TPASS_NCOMPUTE + 32, TPASS_CTAS_PER_SM)
trpass_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TRPassParams p) {
    constexpr uint32_t TILE_BYTES = 1u << TPASS_TILE_BYTES_LOG2;
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full[TPASS_STAGES];
    __shared__ __align__(8) uint64_t done[TPASS_STAGES];

    unsigned char* ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float2* bm = reinterpret_cast<float2*>(ring + TPASS_STAGES * TILE_BYTES);
    uint16_t* tab = reinterpret_cast<uint16_t*>(bm + TRPASS_MAT_ELEMS);

    const int tid = threadIdx.x;
    if (tid == 0) tpass_init_barriers(full, done);
    for (int i = tid; i < p.mat_total; i += blockDim.x) bm[i] = p.m[i];
    if (tid < TPASS_NCOMPUTE) {
        for (int o = 0; o < p.nops; ++o) {
            uint32_t b = 0;
#pragma unroll
            for (int i = 0; i < TPASS_TB; ++i) b ^= (0u - (((uint32_t)tid >> i) & 1u)) & p.lane8[o][i];
            tab[o * TPASS_NCOMPUTE + tid] = (uint16_t)b;
        }
    }
    __syncthreads();

    const uint64_t first = blockIdx.x, stride = gridDim.x;
    const uint64_t nmine = p.ntiles > first ? (p.ntiles - first + stride - 1) / stride : 0;
    if (tid >= TPASS_NCOMPUTE) {
        if (tid == TPASS_NCOMPUTE)
            tpass_producer(&tmap, p.g, p.tp, p.vec_tiles_log2, nmine, first, stride, ring, full, done, (int)sizeof(float2));
        return;
    }
    const TROp op_first = p.op[0];
    const uint32_t b8_first = tab[tid];
    for (uint64_t i = 0; i < nmine; ++i) {
        const int s = (int)(i % TPASS_STAGES);
        unsigned char* tile = ring + s * TILE_BYTES;
        mbar_wait(&full[s], (uint32_t)((i / TPASS_STAGES) & 1));
        TROp op = op_first;
        uint32_t b8 = b8_first;
        for (int o = 0; o < p.nops; ++o) {
            // control of the next op: fetched now, consumed after this op's FMAs
            const int on = o + 1 < p.nops ? o + 1 : 0;
            const TROp op_next = p.op[on];
            const uint32_t b8_next = tab[on * TPASS_NCOMPUTE + tid];
            if (o > 0) {
                if (trop_sync(op) == 2) bar_compute();
                else __syncwarp();
            }
            trtile_thread(tile, b8, op, p.subcode, bm);
            op = op_next;
            b8 = b8_next;
        }
        fence_proxy_async();
        mbar_arrive(&done[s]);
    }
}


// ---- host side ----------------------------------------------------------------------------------
static std::atomic<int64_t> g_tpass_launches{0};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(f);
        else
            (void)cudaGetLastError();
    }
    return fn;
}

// Opt-in (TCB200_TMA=1): on the measured circuits the LDGSTS-staged cpass_kernel is still the
// faster pass (profiles/README.md); the pipeline is correct and covered by the tests.
static bool tma_enabled() {
    const char* e = getenv("TCB200_TMA");
    return e && e[0] == '1';
}

static int sm_count() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
            (void)cudaGetLastError();
            sms = 0;
        }
    }
    return sms;
}

// Tensor map + box plan for tiles of geometry `g` over `batch` state vectors (shared with the gate
// pass, lpass.cu).  > 0: not eligible (alignment, coordinate range, no driver entry point), 0: ready.
int tma_encode_state_map(void* state, int nbits, int is_c64, const TileGeom& g, int64_t batch, TmaPlan* tp, void* map_out) {
    const int apu = is_c64 ? 2 : 1;
    const size_t amp = is_c64 ? 8 : 16;
    if (batch < 1 || (reinterpret_cast<uintptr_t>(state) & 127u)) return 1;
    if (make_tma_plan(g, apu, tp)) return 1;
    const uint64_t total_lines = ((uint64_t)batch << nbits) >> tp->line_bits;
    if (total_lines > 0x7fffffffull) return 1;  // TMA coordinates are int32
    EncodeTiledFn enc = encode_fn();
    if (!enc) return 1;
    const cuuint32_t line_scalars = is_c64 ? 32 : 16;
    cuuint64_t gdim[5] = {line_scalars, total_lines, 1, 1, 1};
    cuuint64_t gstr[4] = {128, 128, 128, 128};
    cuuint32_t box[5] = {line_scalars, (cuuint32_t)1 << (tp->lrow2 - tp->line_bits), 1, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    for (int i = 0; i < tp->nbox_bits; ++i) {
        gdim[2 + i] = 2;
        box[2 + i] = 2;
        gstr[1 + i] = (cuuint64_t)amp << tp->box_bit[i];
    }
    const CUresult cr = enc(static_cast<CUtensorMap*>(map_out), is_c64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5,
                            state, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return cr == CUDA_SUCCESS ? 0 : 1;
}

// Eligibility, tile geometry, box plan and tensor map shared by both kernels.
// >0: not eligible (caller uses the LDGSTS-staged kernels), 0: ready, <0: error
template <typename Real>
static int tpass_prepare(void* state, int nbits, int n_hi, const int* tile_hi, int64_t batch, TileGeom* g, TmaPlan* tp,
                         CUtensorMap* map) {
    using C = typename CT<Real>::type;
    constexpr int APU = CT<Real>::APU;
    constexpr int T = TPASS_TILE_BYTES_LOG2 - (sizeof(Real) == 4 ? 3 : 4);
    if (!tma_enabled()) return 1;
    if (pass_tile_bits(sizeof(Real) == 4 ? TCB200_C64 : TCB200_C128) != T || nbits <= T) return 1;
    if (batch < 1 || (reinterpret_cast<uintptr_t>(state) & 127u)) return 1;
    int rc = make_geom_hi(nbits, T, n_hi, tile_hi, g);
    if (rc) return rc;
    if (make_tma_plan(*g, APU, tp)) return 1;
    const uint64_t total_lines = ((uint64_t)batch << nbits) >> tp->line_bits;
    if (total_lines > 0x7fffffffull) return 1;  // TMA coordinates are int32
    EncodeTiledFn enc = encode_fn();
    if (!enc) return 1;
    // rank-5 tensor map over 128-byte lines (common.cuh TmaPlan)
    const cuuint32_t line_scalars = 128 / sizeof(Real);
    cuuint64_t gdim[5] = {line_scalars, total_lines, 1, 1, 1};
    cuuint64_t gstr[4] = {128, 128, 128, 128};  // bytes, dims 1..4
    cuuint32_t box[5] = {line_scalars, (cuuint32_t)1 << (tp->lrow2 - tp->line_bits), 1, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    for (int i = 0; i < tp->nbox_bits; ++i) {
        gdim[2 + i] = 2;
        box[2 + i] = 2;
        gstr[1 + i] = (cuuint64_t)sizeof(C) << tp->box_bit[i];
    }
    const CUresult cr = enc(map, sizeof(Real) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5,
                            state, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        if (getenv("TCB200_TMA_STRICT")) return fail(TCB200_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);
        return 1;
    }
    return 0;
}

// ---- host plan of a register-tile pass --------------------------------------------------------------
// Every op owns all T local bits of the tile:  4 register bits (the op's target bits + fillers),
// 5 lane bits, and TPASS_TB - 5 warp bits.  Consecutive ops whose target bits leave room for a
// common set of warp bits form a SEGMENT: inside a segment a warp keeps reading and writing the
// same 512 amplitudes, so the ops are separated by __syncwarp() only and the warps of a CTA drift
// apart -- some are loading, some are in the FP32 pipe, some are storing -- instead of marching
// through load / FMA / store phases together.  A CTA barrier is needed only between segments.
constexpr int TR_NWB = TPASS_TB - 5;          // warp-index bits
constexpr int TR_T = TPASS_TILE_BYTES_LOG2 - 3;

// bank class of local amplitude bit b in the SWZ_HW128 layout of complex64 (see abi.cu bank_class)
static int tr_class(int b) { return b == 0 ? 0 : (b <= 6 ? 1 + (b - 1) % 3 : -1); }
static uint32_t tr_swz8(int b) { return swz_amp<2>(1u << b, SWZ_HW128) * (uint32_t)sizeof(float2); }

// candidates outside `taken`, bank-neutral bits (class -1) from the top first, then the others from the top
static int tr_pick(uint32_t taken, bool allow_bit0) {
    for (int b = TR_T - 1; b >= 1; --b)
        if (!((taken >> b) & 1u) && tr_class(b) < 0) return b;
    for (int b = TR_T - 1; b >= 1; --b)
        if (!((taken >> b) & 1u)) return b;
    if (allow_bit0 && !(taken & 1u)) return 0;
    return -1;
}

static int fill_trpass(TRPassParams& q, int nbits, int nrt, const int* rt_k, const int* rt_bits, const int* rt_nsub,
                       const int* sub_k, const int* sub_bits, const double* sub_mats, int64_t batch) {
    if (nrt < 1 || nrt > TCB200_MAX_PASS_OPS) return fail(TCB200_ERR_ARG, "nrt=%d out of range", nrt);
    q.nops = nrt;
    q.nsub = 0;
    // 1. target bits (local) of every op
    uint32_t tmask[TCB200_MAX_PASS_OPS];
    const int* rb = rt_bits;
    for (int o = 0; o < nrt; ++o) {
        const int kt = rt_k[o];
        if (kt < 1 || kt > 4) return fail(TCB200_ERR_UNSUPPORTED, "register tile of %d bits (max 4)", kt);
        tmask[o] = 0;
        for (int i = 0; i < kt; ++i) {
            if (rb[i] < 0 || rb[i] >= nbits) return fail(TCB200_ERR_ARG, "bit %d out of range", rb[i]);
            const int lb = local_bit(q.g, rb[i]);
            if (lb < 0) return fail(TCB200_ERR_ARG, "bit %d is not inside the tile", rb[i]);
            if ((tmask[o] >> lb) & 1u) return fail(TCB200_ERR_ARG, "bit %d repeated in a register tile", rb[i]);
            tmask[o] |= 1u << lb;
        }
        rb += kt;
    }
    // 2. segments and their warp bits
    uint32_t wmask[TCB200_MAX_PASS_OPS];
    int wbits[TCB200_MAX_PASS_OPS][8];
    int sync_kind[TCB200_MAX_PASS_OPS];
    for (int o0 = 0; o0 < nrt;) {
        uint32_t u = tmask[o0];
        int o1 = o0 + 1;
        while (o1 < nrt && __builtin_popcount(u | tmask[o1]) <= TR_T - TR_NWB) u |= tmask[o1++];
        uint32_t w = 0;
        int wb[8];
        for (int i = 0; i < TR_NWB; ++i) {
            // bit 0 stays free for the 16-byte accesses whenever there is a choice
            wb[i] = tr_pick(u | w, true);
            if (wb[i] < 0) return fail(TCB200_ERR_UNSUPPORTED, "no room for the warp bits of a segment");
            w |= 1u << wb[i];
        }
        for (int o = o0; o < o1; ++o) {
            wmask[o] = w;
            for (int i = 0; i < TR_NWB; ++i) wbits[o][i] = wb[i];
            sync_kind[o] = (o == o0) ? 2 : 1;
        }
        o0 = o1;
    }
    // 3. per op: fillers, register-bit order, lane order, gate positions, matrices
    int moff = 0;
    rb = rt_bits;
    const int* sk = sub_k;
    const int* sb = sub_bits;
    const double* mp = sub_mats;
    for (int o = 0; o < nrt; ++o) {
        const int kt = rt_k[o];
        uint32_t r = tmask[o];
        if (__builtin_popcount(r) < 4 && !((r | wmask[o]) & 1u)) r |= 1u;
        while (__builtin_popcount(r) < 4) {
            const int b = tr_pick(r | wmask[o], true);
            if (b < 0) return fail(TCB200_ERR_UNSUPPORTED, "no room for the filler bits of a register tile");
            r |= 1u << b;
        }
        int reg[4], nr = 0;
        for (int b = 0; b < TR_T; ++b)
            if ((r >> b) & 1u) reg[nr++] = b;
        TROp& op = q.op[o];
        const int vec0 = reg[0] == 0 ? 1 : 0;
        if (rt_nsub[o] < 1 || rt_nsub[o] > 255) return fail(TCB200_ERR_ARG, "register tile with %d gates", rt_nsub[o]);
        op.t01 = tr_swz8(reg[0]) | (tr_swz8(reg[1]) << 16);
        op.t23 = tr_swz8(reg[2]) | (tr_swz8(reg[3]) << 16);
        op.flags = (uint32_t)rt_nsub[o] | ((uint32_t)vec0 << 8) | ((uint32_t)sync_kind[o] << 12) | ((uint32_t)q.nsub << 16);
        const int first_sub = q.nsub;
        // lanes: one bit per bank class first (16-byte accesses need classes 1..3 on lanes 0..2,
        // 8-byte accesses class 0 as well on lanes 0..3), then the rest ascending
        const uint32_t lanes = ((1u << TR_T) - 1u) & ~(r | wmask[o]);
        int order[16], no = 0;
        uint32_t used = 0;
        for (int c = vec0 ? 1 : 0; c < 4; ++c)
            for (int b = 0; b < TR_T; ++b)
                if (((lanes >> b) & 1u) && !((used >> b) & 1u) && tr_class(b) == c) {
                    order[no++] = b;
                    used |= 1u << b;
                    break;
                }
        for (int b = 0; b < TR_T; ++b)
            if (((lanes >> b) & 1u) && !((used >> b) & 1u)) order[no++] = b;
        if (no != 5) return fail(TCB200_ERR_UNSUPPORTED, "unexpected tile size for the register-tile pass");
        for (int i = 0; i < 5; ++i) q.lane8[o][i] = tr_swz8(order[i]);
        for (int i = 0; i < TR_NWB; ++i) q.lane8[o][5 + i] = tr_swz8(wbits[o][i]);
        for (int s = 0; s < rt_nsub[o]; ++s) {
            if (q.nsub >= TRPASS_MAX_SUB) return fail(TCB200_ERR_UNSUPPORTED, "more than %d gates in one pass", TRPASS_MAX_SUB);
            const int k = *sk++;
            if (k != 1 && k != 2) return fail(TCB200_ERR_UNSUPPORTED, "gate of %d bits inside a register tile", k);
            int pos[2] = {0, 0};
            for (int i = 0; i < k; ++i) {
                bool in_tile = false;
                for (int j = 0; j < kt; ++j) in_tile = in_tile || rb[j] == sb[i];
                if (!in_tile) return fail(TCB200_ERR_ARG, "gate bit %d is not in its register tile", sb[i]);
                const int lb = local_bit(q.g, sb[i]);
                for (int j = 0; j < 4; ++j)
                    if (reg[j] == lb) pos[i] = j;
            }
            if (k == 2 && pos[1] <= pos[0]) return fail(TCB200_ERR_ARG, "gate bits must be ascending");
            // every gate owns a 16-element slot (1-bit gates use the first 4): the kernel always
            // fetches a whole slot before it looks at the gate code
            if (moff + 16 > TRPASS_MAT_ELEMS) return fail(TCB200_ERR_UNSUPPORTED, "more than %d gates in one pass", TRPASS_MAT_ELEMS / 16);
            static const int pair_case[4][4] = {{-1, 0, 1, 2}, {-1, -1, 3, 4}, {-1, -1, -1, 5}, {-1, -1, -1, -1}};
            const int cs = k == 2 ? pair_case[pos[0]][pos[1]] : 6 + pos[0];
            q.subcode[q.nsub++] = (uint32_t)cs | ((uint32_t)moff << 8);
            const int sz = 1 << (2 * k);
            for (int i = 0; i < 16; ++i)
                q.m[moff + i] = i < sz ? make_float2((float)mp[2 * i], (float)mp[2 * i + 1]) : make_float2(0.f, 0.f);
            moff += 16;
            mp += 2 * sz;
            sb += k;
        }
        op.code0 = q.subcode[first_sub];
        rb += kt;
    }
    q.mat_total = moff;
    q.vec_tiles_log2 = nbits - TR_T;
    q.ntiles = (uint64_t)batch << (nbits - TR_T);
    return 0;
}

constexpr size_t TRPASS_SMEM = 1024 + (size_t)TPASS_STAGES * (1u << TPASS_TILE_BYTES_LOG2) + TRPASS_MAT_ELEMS * sizeof(float2) +
                               (size_t)TCB200_MAX_PASS_OPS * TPASS_NCOMPUTE * sizeof(uint16_t);

// complex64 register-tile pass through the TMA pipeline; same return convention as launch_tpass
int launch_trpass(void* state, int nbits, int nrt, const int* rt_k, const int* rt_bits, const int* rt_nsub,
                  const int* sub_k, const int* sub_bits, const double* sub_mats, int n_hi, const int* tile_hi,
                  int64_t batch, cudaStream_t st) {
    static thread_local TRPassParams* tp = nullptr;
    if (!tp) tp = new TRPassParams();
    TRPassParams& q = *tp;
    alignas(64) CUtensorMap map;
    int rc = tpass_prepare<float>(state, nbits, n_hi, tile_hi, batch, &q.g, &q.tp, &map);
    if (rc) return rc;
    rc = fill_trpass(q, nbits, nrt, rt_k, rt_bits, rt_nsub, sub_k, sub_bits, sub_mats, batch);
    if (rc) return rc;
    static bool attr = false;
    if (!attr) {
        TCB_CUDA(cudaFuncSetAttribute(trpass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRPASS_SMEM));
        attr = true;
    }
    const int sms = sm_count();
    if (sms <= 0) return fail(TCB200_ERR_CUDA, "cannot query the SM count");
    const uint64_t ctas = (uint64_t)sms * TPASS_CTAS_PER_SM;
    const unsigned grid = (unsigned)(q.ntiles < ctas ? q.ntiles : ctas);
    trpass_kernel<<<grid, TPASS_NCOMPUTE + 32, TRPASS_SMEM, st>>>(map, q);
    TCB_LAUNCH_CHECK("trpass_kernel");
    g_tpass_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

// >0: not eligible (caller uses cpass_kernel), 0: launched, <0: error
template <typename Real>
static int launch_tpass_t(void* state, int nbits, int nops, const int* ops_k, const int* ops_bits,
                          const double* mats, int n_hi, const int* tile_hi, int64_t batch, cudaStream_t st) {
    using C = typename CT<Real>::type;
    constexpr int APU = CT<Real>::APU;
    constexpr int T = TPASS_TILE_BYTES_LOG2 - (sizeof(Real) == 4 ? 3 : 4);
    constexpr int MAXM = TPASS_MAT_BYTES / (int)sizeof(C);
    if (sizeof(Real) == 4) {
        // complex64 passes made of 1-/2-bit blocks only: the register-tile kernel, one gate per tile
        bool narrow = true;
        for (int o = 0; o < nops; ++o) narrow = narrow && ops_k[o] <= 2;
        if (narrow) {
            int one[TCB200_MAX_PASS_OPS];
            for (int o = 0; o < nops; ++o) one[o] = 1;
            return launch_trpass(state, nbits, nops, ops_k, ops_bits, one, ops_k, ops_bits, mats, n_hi, tile_hi, batch, st);
        }
    }
    static thread_local TPassParams<Real>* tp = nullptr;
    if (!tp) tp = new TPassParams<Real>();
    TPassParams<Real>& q = *tp;
    alignas(64) CUtensorMap map;
    int rc = tpass_prepare<Real>(state, nbits, n_hi, tile_hi, batch, &q.g, &q.tp, &map);
    if (rc) return rc;

    q.nops = nops;
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
        if (moff + sz > MAXM) return fail(TCB200_ERR_UNSUPPORTED, "pass matrices exceed %d bytes", TPASS_MAT_BYTES);
        q.op[o].k = k;
        q.op[o].moff = moff;
        rc = make_group_map(q.g, APU, k, b, &q.op[o].gm, SWZ_HW128);
        if (rc) return rc;
        for (int i = 0; i < sz; ++i) {
            q.m[moff + i].x = (Real)mp[2 * i];
            q.m[moff + i].y = (Real)mp[2 * i + 1];
        }
        moff += sz;
        mp += 2 * sz;
        b += k;
    }
    q.mat_total = moff;
    q.vec_tiles_log2 = nbits - T;
    q.ntiles = (uint64_t)batch << (nbits - T);

    static bool attr = false;
    const size_t smem = 1024 + (size_t)TPASS_STAGES * (1u << TPASS_TILE_BYTES_LOG2) + TPASS_MAT_BYTES;
    if (!attr) {
        TCB_CUDA(cudaFuncSetAttribute(tpass_kernel<Real>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    const int sms = sm_count();
    if (sms <= 0) return fail(TCB200_ERR_CUDA, "cannot query the SM count");
    const uint64_t ctas = (uint64_t)sms * TPASS_CTAS_PER_SM;
    const unsigned grid = (unsigned)(q.ntiles < ctas ? q.ntiles : ctas);
    tpass_kernel<Real><<<grid, TPASS_NCOMPUTE + 32, smem, st>>>(map, q);
    TCB_LAUNCH_CHECK("tpass_kernel");
    g_tpass_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

int launch_tpass(int dtype, void* state, int nbits, int nops, const int* ops_k, const int* ops_bits,
                 const double* mats, int n_hi, const int* tile_hi, int64_t batch, cudaStream_t st) {
    if (dtype == TCB200_C64)
        return launch_tpass_t<float>(state, nbits, nops, ops_k, ops_bits, mats, n_hi, tile_hi, batch, st);
    return launch_tpass_t<double>(state, nbits, nops, ops_k, ops_bits, mats, n_hi, tile_hi, batch, st);
}

}  // namespace tcb

extern "C" int64_t tcb200_tma_pass_count(void) { return tcb::g_tpass_launches.load(); }

namespace tcb {

#ifdef TCB200_EMU
// tests/emu only: the compute half of tpass_kernel on the CPU, with a software model of the
// tensor-map boxes (dense box order d4..d0, SWIZZLE_128B on the destination address) in place
// of the TMA unit.
template <typename Real>
static int emu_tpass_t(void* state, int nbits, int nops, const int* ops_k, const int* ops_bits,
                       const double* mats, int n_hi, const int* tile_hi, int64_t batch) {
    using C = typename CT<Real>::type;
    constexpr int APU = CT<Real>::APU;
    constexpr int T = TPASS_TILE_BYTES_LOG2 - (sizeof(Real) == 4 ? 3 : 4);
    if (nbits <= T) return 1;
    TileGeom g;
    TmaPlan tp;
    int rc = make_geom_hi(nbits, T, n_hi, tile_hi, &g);
    if (rc) return rc;
    if (make_tma_plan(g, APU, &tp)) return 1;
    std::vector<PassOp> ops(nops);
    std::vector<C> ms;
    const int* b = ops_bits;
    const double* mp = mats;
    for (int o = 0; o < nops; ++o) {
        const int k = ops_k[o], sz = 1 << (2 * k);
        ops[o].k = k;
        ops[o].moff = (int)ms.size();
        rc = make_group_map(g, APU, k, b, &ops[o].gm, SWZ_HW128);
        if (rc) return rc;
        for (int i = 0; i < sz; ++i) ms.push_back(mk<C, Real>((Real)mp[2 * i], (Real)mp[2 * i + 1]));
        b += k;
        mp += 2 * sz;
    }
    std::vector<C> tilev((size_t)1 << T);
    C* tile = tilev.data();
    C* vec = static_cast<C*>(state);
    const uint64_t ntiles = (uint64_t)batch << (nbits - T);
    const uint64_t vmask = (1ull << (nbits - T)) - 1ull;
    auto boxes = [&](uint64_t base, bool store) {
        for (uint32_t c = 0; c < (1u << tp.nextra); ++c) {
            const uint64_t line0 = (base + tma_box_offset(tp, c)) >> tp.line_bits;  // coordinate d1
            uint32_t dense = (uint32_t)c << tp.box_amps_log2;                       // amplitudes, box order
            for (uint32_t hi = 0; hi < (1u << tp.nbox_bits); ++hi) {
                uint64_t off = 0;
                for (int i = 0; i < tp.nbox_bits; ++i)
                    if ((hi >> i) & 1u) off += 1ull << tp.box_bit[i];
                for (uint32_t e = 0; e < (1u << tp.lrow2); ++e, ++dense) {
                    const uint64_t gi = (line0 << tp.line_bits) + off + e;
                    const uint32_t byte = dense * (uint32_t)sizeof(C);
                    const uint32_t sw = byte ^ (((byte >> 7) & 7u) << 4);
                    C* sp = reinterpret_cast<C*>(reinterpret_cast<unsigned char*>(tile) + sw);
                    if (store) vec[gi] = *sp;
                    else *sp = vec[gi];
                }
            }
        }
    };
    for (uint64_t t = 0; t < ntiles; ++t) {
        const uint64_t base = ((t >> (nbits - T)) << nbits) + tile_base(g, t & vmask);
        boxes(base, false);
        for (int o = 0; o < nops; ++o) {
            const C* m = ms.data() + ops[o].moff;
            for (int tid = 0; tid < TPASS_NCOMPUTE; ++tid) {
                switch (ops[o].k) {
                    case 1: apply_block_tb<C, 1, T, TPASS_TB>(tile, ops[o].gm, tid, [&](int i, int j) { return m[i * 2 + j]; }); break;
                    case 2: apply_block_tb<C, 2, T, TPASS_TB>(tile, ops[o].gm, tid, [&](int i, int j) { return m[i * 4 + j]; }); break;
                    case 3: apply_block_tb<C, 3, T, TPASS_TB>(tile, ops[o].gm, tid, [&](int i, int j) { return m[i * 8 + j]; }); break;
                    default: apply_block_tb<C, 4, T, TPASS_TB>(tile, ops[o].gm, tid, [&](int i, int j) { return m[i * 16 + j]; }); break;
                }
            }
        }
        boxes(base, true);
    }
    return 0;
}

// trpass_kernel on the CPU: same parameter block (fill_trpass), same per-thread body, same box model
extern "C" __attribute__((visibility("default"))) int emu_apply_trpass(void* state, int nbits, int nrt, const int* rt_k,
                                                                        const int* rt_bits, const int* rt_nsub, const int* sub_k,
                                                                        const int* sub_bits, const double* sub_mats, int n_hi,
                                                                        const int* tile_hi, int64_t batch) {
    constexpr int T = TPASS_TILE_BYTES_LOG2 - 3;
    if (nbits <= T) return 1;
    std::vector<TRPassParams> qv(1);
    TRPassParams& q = qv[0];
    int rc = make_geom_hi(nbits, T, n_hi, tile_hi, &q.g);
    if (rc) return rc;
    if (make_tma_plan(q.g, 2, &q.tp)) return 1;
    rc = fill_trpass(q, nbits, nrt, rt_k, rt_bits, rt_nsub, sub_k, sub_bits, sub_mats, batch);
    if (rc) return rc;
    const TmaPlan& tp = q.tp;
    std::vector<float2> tilev((size_t)1 << T);
    unsigned char* tile = reinterpret_cast<unsigned char*>(tilev.data());
    float2* vec = static_cast<float2*>(state);
    const uint64_t vmask = (1ull << (nbits - T)) - 1ull;
    auto boxes = [&](uint64_t base, bool store) {
        for (uint32_t c = 0; c < (1u << tp.nextra); ++c) {
            const uint64_t line0 = (base + tma_box_offset(tp, c)) >> tp.line_bits;
            uint32_t dense = (uint32_t)c << tp.box_amps_log2;
            for (uint32_t hi = 0; hi < (1u << tp.nbox_bits); ++hi) {
                uint64_t off = 0;
                for (int i = 0; i < tp.nbox_bits; ++i)
                    if ((hi >> i) & 1u) off += 1ull << tp.box_bit[i];
                for (uint32_t e = 0; e < (1u << tp.lrow2); ++e, ++dense) {
                    const uint64_t gi = (line0 << tp.line_bits) + off + e;
                    const uint32_t byte = dense * (uint32_t)sizeof(float2);
                    float2* sp = reinterpret_cast<float2*>(tile + (byte ^ (((byte >> 7) & 7u) << 4)));
                    if (store) vec[gi] = *sp;
                    else *sp = vec[gi];
                }
            }
        }
    };
    std::vector<uint16_t> tab((size_t)q.nops * TPASS_NCOMPUTE);
    for (int o = 0; o < q.nops; ++o)
        for (int tid = 0; tid < TPASS_NCOMPUTE; ++tid) {
            uint32_t b = 0;
            for (int i = 0; i < TPASS_TB; ++i) b ^= (0u - (((uint32_t)tid >> i) & 1u)) & q.lane8[o][i];
            if (b >= (1u << TPASS_TILE_BYTES_LOG2)) return fail(TCB200_ERR_ARG, "group offset does not fit 16 bits");
            tab[(size_t)o * TPASS_NCOMPUTE + tid] = (uint16_t)b;
        }
    // The sequential emulation cannot see a missing barrier, so check the plan itself: an op that
    // is separated from its predecessor by __syncwarp() only (sync == 1) must touch, warp by warp,
    // exactly the amplitudes that warp touched in the predecessor.
    {
        std::vector<std::vector<uint32_t>> prev;
        for (int o = 0; o < q.nops; ++o) {
            std::vector<std::vector<uint32_t>> cur(TPASS_NCOMPUTE / 32);
            for (int tid = 0; tid < TPASS_NCOMPUTE; ++tid)
                for (uint32_t j = 0; j < 16; ++j) {
                    uint32_t a = tab[(size_t)o * TPASS_NCOMPUTE + tid];
                    for (int i = 0; i < 4; ++i)
                        if ((j >> i) & 1u) a ^= trop_t8(q.op[o], i);
                    cur[tid / 32].push_back(a);
                }
            for (auto& w : cur) std::sort(w.begin(), w.end());
            if (o > 0 && trop_sync(q.op[o]) != 2 && cur != prev)
                return fail(TCB200_ERR_ARG, "op %d is not warp-private with respect to op %d", o, o - 1);
            if (o > 0 && trop_sync(q.op[o]) != 1 && trop_sync(q.op[o]) != 2) return fail(TCB200_ERR_ARG, "op %d has no sync kind", o);
            prev.swap(cur);
        }
    }
    for (uint64_t t = 0; t < q.ntiles; ++t) {
        const uint64_t base = ((t >> (nbits - T)) << nbits) + tile_base(q.g, t & vmask);
        boxes(base, false);
        for (int o = 0; o < q.nops; ++o)
            for (int tid = 0; tid < TPASS_NCOMPUTE; ++tid)
                trtile_thread(tile, tab[(size_t)o * TPASS_NCOMPUTE + tid], q.op[o], q.subcode, q.m);
        boxes(base, true);
    }
    return 0;
}

extern "C" __attribute__((visibility("default"))) int emu_apply_tpass(void* state, int nbits, int dtype, int nops,
                                                                       const int* ops_k, const int* ops_bits, const double* mats,
                                                                       int n_hi, const int* tile_hi, int64_t batch) {
    if (dtype == TCB200_C64) return emu_tpass_t<float>(state, nbits, nops, ops_k, ops_bits, mats, n_hi, tile_hi, batch);
    return emu_tpass_t<double>(state, nbits, nops, ops_k, ops_bits, mats, n_hi, tile_hi, batch);
}

// worst bank-conflict degree of one block's shared-memory phases under the SWZ_HW128 layout
extern "C" __attribute__((visibility("default"))) int emu_conflict_degree_tpass(int nbits, int dtype, int k, const int* bits,
                                                                                 int n_hi, const int* tile_hi) {
    TileGeom g;
    GroupMap gm;
    const int apu = dtype == TCB200_C64 ? 2 : 1;
    const int T = TPASS_TILE_BYTES_LOG2 - (dtype == TCB200_C64 ? 3 : 4);
    if (make_geom_hi(nbits, T, n_hi, tile_hi, &g)) return -1;
    if (make_group_map(g, apu, k, bits, &gm, SWZ_HW128)) return -1;
    const int esz = dtype == TCB200_C64 ? 8 : 16;
    const bool vec = gm.vec0 || esz == 16;
    const int lanes_per_phase = vec ? 8 : 16;
    int worst = 1;
    const uint32_t ngroups = 1u << gm.ngb;
    for (uint32_t w0 = 0; w0 < (uint32_t)TPASS_NCOMPUTE && w0 < ngroups; w0 += lanes_per_phase) {
        for (uint32_t j = 0; j < (1u << k); ++j) {
            int cnt[32] = {0};
            for (int l = 0; l < lanes_per_phase; ++l) {
                const uint32_t gi = w0 + l;
                if (gi >= ngroups) break;
                const uint32_t byte = (group_base(gm, gi) ^ gm.tval[j]) * esz;
                cnt[vec ? (byte / 16) % 8 : (byte / 8) % 16]++;
            }
            for (int b = 0; b < 32; ++b)
                if (cnt[b] > worst) worst = cnt[b];
        }
    }
    return worst;
}
#endif

}  // namespace tcb
