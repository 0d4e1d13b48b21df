// Shared declarations for the tcb200 engine (sm_100a).
//
// Everything that computes an address or an index is __host__ __device__ so that
// tests/emu/ can execute the exact kernel bodies thread by thread on the CPU (the build
// container has no GPU); the product never runs those host instantiations.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "tcb200.h"

#define TCB_HD __host__ __device__ __forceinline__

namespace tcb {

// ---- error plumbing (abi.cu) ---------------------------------------------------------------
int fail(int code, const char* fmt, ...);
void count_launch(int n = 1);
int cuda_fail(cudaError_t e, const char* what);

#define TCB_CUDA(x)                                          \
    do {                                                     \
        cudaError_t _e = (x);                                \
        if (_e != cudaSuccess) return cuda_fail(_e, #x);     \
    } while (0)

#define TCB_LAUNCH_CHECK(name)                               \
    do {                                                     \
        cudaError_t _e = cudaGetLastError();                 \
        if (_e != cudaSuccess) return cuda_fail(_e, name);   \
        count_launch();                                      \
    } while (0)

// ---- complex types -------------------------------------------------------------------------
template <typename Real>
struct CT;
template <>
struct CT<float> {
    using type = float2;
    static constexpr int APU = 2;  // amplitudes per 16-byte unit
};
template <>
struct CT<double> {
    using type = double2;
    static constexpr int APU = 1;
};

template <typename C, typename R>
TCB_HD C mk(R x, R y) {
    C c;
    c.x = x;
    c.y = y;
    return c;
}

// acc += m * v  (4 FMA)
template <typename C>
TCB_HD void cfma(C& acc, const C m, const C v) {
    acc.x = fma(m.x, v.x, acc.x);
    acc.x = fma(-m.y, v.y, acc.x);
    acc.y = fma(m.x, v.y, acc.y);
    acc.y = fma(m.y, v.x, acc.y);
}

template <typename C>
TCB_HD C cmul(const C a, const C b) {
    C r;
    r.x = a.x * b.x - a.y * b.y;
    r.y = a.x * b.y + a.y * b.x;
    return r;
}

// complex64 on Blackwell: one complex multiply-add = 2 packed FFMA2 (sm_100 fma.rn.f32x2).
// ptxas folds the broadcast of m.x / m.y, the (v.y, v.x) half swap and the (-,+) sign pattern
// into operand modifiers (UR.F32 scalar operand, .F32x2.LO_HI, .NP), so the matrix stays a plain
// (re, im) pair in the constant bank and the instruction count per amplitude halves -- which
// is what moves the k=4 block from issue-bound to HBM-bound (profiles/r1_dense_k4.md).
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
template <>
__device__ __forceinline__ void cfma<float2>(float2& acc, const float2 m, const float2 v) {
    acc = __ffma2_rn(make_float2(m.x, m.x), v, acc);
    acc = __ffma2_rn(make_float2(-m.y, m.y), make_float2(v.y, v.x), acc);
}
template <>
__device__ __forceinline__ float2 cmul<float2>(const float2 m, const float2 v) {
    float2 acc = __ffma2_rn(make_float2(m.x, m.x), v, make_float2(0.f, 0.f));
    return __ffma2_rn(make_float2(-m.y, m.y), make_float2(v.y, v.x), acc);
}
#endif

// Row accumulator  out = sum_j m_j * v_j.  The generic version chains cmul / cfma.  For complex64
// on sm_100 the two packed products are kept apart,
//     P += (m.x, m.x) * (v.x, v.y)        Q += (m.y, m.y) * (v.x, v.y)
//     out = (P.x - Q.y, P.y + Q.x) = (-1, 1) * (Q.y, Q.x) + P        (one more FFMA2 per row)
// because then every FFMA2 takes the matrix element as a single broadcast register (R.F32) and
// the amplitude pair exactly as loaded.  The cfma<float2> form above needs the register PAIR
// (m.y, m.y) for its (-,+) operand; with the matrix in registers (cpass / tpass) ptxas rebuilds
// that pair with two MOVs for almost every FFMA2 (measured: 69 % of the issued instructions of a
// 2-bit block were not FFMA2).  The swap and the sign of the final step fold into operand
// modifiers (.LO_HI, .NP) and a uniform-register constant.
template <typename C>
struct RowAcc {
    C a;
    TCB_HD void init(const C m, const C v) { a = cmul(m, v); }
    TCB_HD void mac(const C m, const C v) { cfma(a, m, v); }
    TCB_HD C result() const { return a; }
};
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
template <>
struct RowAcc<float2> {
    float2 p, q;
    __device__ __forceinline__ void init(const float2 m, const float2 v) {
        p = __fmul2_rn(make_float2(m.x, m.x), v);
        q = __fmul2_rn(make_float2(m.y, m.y), v);
    }
    __device__ __forceinline__ void mac(const float2 m, const float2 v) {
        p = __ffma2_rn(make_float2(m.x, m.x), v, p);
        q = __ffma2_rn(make_float2(m.y, m.y), v, q);
    }
    __device__ __forceinline__ float2 result() const {
        return __ffma2_rn(make_float2(-1.f, 1.f), make_float2(q.y, q.x), p);
    }
};
#endif

// The chained form (cmul / cfma, 2 FFMA2 per complex multiply-add, no final step) stays the
// choice where the FP32 pipe itself is the limit and the MOVs hide in spare issue slots
// (cpass_kernel: 10.25 ms against 10.97 ms for a 12-block pass at n = 30).
template <typename C>
struct ChainAcc {
    C a;
    TCB_HD void init(const C m, const C v) { a = cmul(m, v); }
    TCB_HD void mac(const C m, const C v) { cfma(a, m, v); }
    TCB_HD C result() const { return a; }
};
template <typename C, bool SPLIT>
struct AccSel {
    using type = ChainAcc<C>;
};
template <typename C>
struct AccSel<C, true> {
    using type = RowAcc<C>;
};

// ---- shared-memory swizzle -----------------------------------------------------------------
// Tiles live in shared memory as 16-byte units.  Unit u is stored at slot
//   swz(u) = u ^ ((u>>3)&7) ^ ((u>>6)&7)
// i.e. the 3 slot bits that select the 16-byte bank group are XORed with the next two 3-bit
// fields.  The map is GF(2)-linear and only touches the low 3 bits, so
//   swz(a ^ b) == swz(a) ^ swz(b)   and   it permutes every aligned run of 8 units.
// Consequences: (i) staging a tile with consecutive lanes -> consecutive units is
// conflict-free; (ii) a gate group can be addressed as swz(base) ^ swz(target offset), both
// precomputed; (iii) the host can order the lane bits of the group index so that the lanes of
// a quarter/half warp always hit distinct banks, whatever the target bits are.
TCB_HD uint32_t swz_unit(uint32_t u) { return u ^ ((u >> 3) & 7u) ^ ((u >> 6) & 7u); }

// Layout modes of a staged tile:
//   SWZ_NONE  linear;
//   SWZ_SW    the two-term software swizzle above (tiles staged by LDGSTS);
//   SWZ_HW128 the TMA hardware pattern CU_TENSOR_MAP_SWIZZLE_128B: inside every 1024-byte block the
//             16-byte chunk index (address bits 4..6) is XORed with the 128-byte line index
//             (address bits 7..9), i.e. only the first term.  Tiles written by
//             cp.async.bulk.tensor (tpass_kernel) arrive in this layout.
constexpr int SWZ_NONE = 0, SWZ_SW = 1, SWZ_HW128 = 2;

// bytes (log2) of a multi-block pass tile: 8192 complex64 / 4096 complex128 amplitudes.  15 selects
// the two-rings-per-SM build of the TMA pipeline (tpass.cu); measurements of both in DESIGN.md.
constexpr int PASS_TILE_BYTES_LOG2 = 16;

TCB_HD uint32_t swz_unit_m(int mode, uint32_t u) {
    if (mode == SWZ_SW) return swz_unit(u);
    if (mode == SWZ_HW128) return u ^ ((u >> 3) & 7u);
    return u;
}

template <int APU>
TCB_HD uint32_t swz_amp(uint32_t e, int mode = SWZ_SW) {
    if (APU == 2) return (swz_unit_m(mode, e >> 1) << 1) | (e & 1u);
    return swz_unit_m(mode, e);
}

// ---- tile geometry -------------------------------------------------------------------------
// A tile is the set of 2^T amplitudes whose index agrees outside the bit set
//   S = [0, lrow)  U  { hb[0] < hb[1] < ... < hb[h-1] },  hb[j] >= lrow,  T = lrow + h.
// Local index e (T bits): low lrow bits = position inside a contiguous row, bit lrow+j = hb[j].
struct TileGeom {
    int n;      // bits of one state vector
    int T;      // log2(tile amplitudes)
    int h;      // gathered high bits
    int lrow;   // T - h: log2(row length)
    int hb[10];  // ascending (the kernels with a shared-memory row table take h <= 8; the gate pass's
                 // production shape addresses rows through host constants and takes 9)
};

TCB_HD uint64_t tile_base(const TileGeom& g, uint64_t t) {
    uint64_t x = t << g.lrow;
    for (int j = 0; j < g.h; ++j) {
        const int pos = g.hb[j];
        const uint64_t lo = x & ((1ull << pos) - 1ull);
        x = ((x >> pos) << (pos + 1)) | lo;
    }
    return x;
}

TCB_HD uint64_t row_offset(const TileGeom& g, uint32_t row) {
    uint64_t o = 0;
    for (int j = 0; j < g.h; ++j)
        if ((row >> j) & 1u) o |= 1ull << g.hb[j];
    return o;
}

// global bit position -> local tile bit (or -1)
inline int local_bit(const TileGeom& g, int b) {
    if (b < g.lrow) return b;
    for (int j = 0; j < g.h; ++j)
        if (g.hb[j] == b) return g.lrow + j;
    return -1;
}

// deposit the low bits of j into the positions pos[0..k)
TCB_HD uint32_t deposit(uint32_t j, const int* pos, int k) {
    uint32_t o = 0;
    for (int i = 0; i < k; ++i)
        if ((j >> i) & 1u) o |= 1u << pos[i];
    return o;
}

// ---- staging (global <-> shared), 16 bytes per lane ------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
#if defined(__CUDA_ARCH__)
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_wait_all() {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
#endif
}

__device__ __forceinline__ void cp_async_commit() {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.commit_group;\n" ::: "memory");
#endif
}
template <int N>
__device__ __forceinline__ void cp_async_wait_group() {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
#endif
}

struct alignas(16) Unit16 {
    uint32_t w[4];
};

// Bring the tile with base amplitude index `base` into shared memory.  `rowoff` holds
// row_offset() of the 2^h rows.  SWZ selects the layout (SWZ_NONE / SWZ_SW / SWZ_HW128).
template <typename C, int SWZ>
TCB_HD void stage_in(const TileGeom& g, const C* vec, uint64_t base, C* tile,
                     const uint64_t* rowoff, int tid, int nthr) {
    constexpr int APU = 16 / (int)sizeof(C);
    const uint32_t nunits = (1u << g.T) / APU;
    const uint32_t rowmask = (1u << g.lrow) - 1u;
    Unit16* t16 = reinterpret_cast<Unit16*>(tile);
    for (uint32_t u = tid; u < nunits; u += nthr) {
        const uint32_t e = u * APU;
        const uint64_t gi = base + rowoff[e >> g.lrow] + (e & rowmask);
        const uint32_t slot = swz_unit_m(SWZ, u);
#if defined(__CUDA_ARCH__)
        cp_async16(t16 + slot, vec + gi);
#else
        t16[slot] = *reinterpret_cast<const Unit16*>(vec + gi);
#endif
    }
}

template <typename C, int SWZ>
TCB_HD void stage_out(const TileGeom& g, C* vec, uint64_t base, const C* tile,
                      const uint64_t* rowoff, int tid, int nthr) {
    constexpr int APU = 16 / (int)sizeof(C);
    const uint32_t nunits = (1u << g.T) / APU;
    const uint32_t rowmask = (1u << g.lrow) - 1u;
    const Unit16* t16 = reinterpret_cast<const Unit16*>(tile);
    for (uint32_t u = tid; u < nunits; u += nthr) {
        const uint32_t e = u * APU;
        const uint64_t gi = base + rowoff[e >> g.lrow] + (e & rowmask);
        const uint32_t slot = swz_unit_m(SWZ, u);
        *reinterpret_cast<Unit16*>(vec + gi) = t16[slot];
    }
}

// ---- dense block on a staged tile ----------------------------------------------------------
// Group index gidx (T-K bits) -> swizzled local amplitude index of the group's element 0:
// XOR of ntval[i] over the set bits i of gidx.  The element with target combination j is at
// (that) ^ tval[j].
struct GroupMap {
    int ngb;              // T - K
    uint32_t ntval[16];   // swz_amp(1 << (local non-target bit i)), in host-chosen lane order
    uint32_t tval[32];    // swz_amp(deposit(j, local target bits))
    int vec0;             // 1: local bit 0 is target bit 0 and sizeof(C)==8 -> pairs are one 16B unit
};

// one fused block of a multi-block pass
struct PassOp {
    int k;
    int moff;  // offset (in complex elements) of this block's matrix inside the pass blob
    GroupMap gm;
};

TCB_HD uint32_t group_base(const GroupMap& m, uint32_t gidx) {
    uint32_t b = 0;
    for (int i = 0; i < m.ngb; ++i) b ^= (0u - ((gidx >> i) & 1u)) & m.ntval[i];
    return b;
}

// One group: v <- M v with M(i, j) supplied by `mat` (any callable returning C).
template <typename C, int K, typename Mat, bool SPLIT = false>
TCB_HD void apply_group(C* tile, const uint32_t base, const uint32_t* tval, const bool vec0,
                        const Mat& mat) {
    constexpr int D = 1 << K;
    using Acc = typename AccSel<C, SPLIT>::type;
    C v[D];
    if (sizeof(C) == 8 && vec0) {
        // bit 0 of the local index is target bit 0: (j, j|1) share one aligned 16-byte unit
#pragma unroll
        for (int j = 0; j < D; j += 2) {
            const Unit16 q = *reinterpret_cast<const Unit16*>(tile + (base ^ tval[j]));
            const C* qc = reinterpret_cast<const C*>(&q);
            v[j] = qc[0];
            v[j + 1] = qc[1];
        }
    } else {
#pragma unroll
        for (int j = 0; j < D; ++j) v[j] = tile[base ^ tval[j]];
    }
    if (sizeof(C) == 8 && vec0) {
#pragma unroll
        for (int i = 0; i < D; i += 2) {
            Unit16 q;
            C* qc = reinterpret_cast<C*>(&q);
#pragma unroll
            for (int ii = 0; ii < 2; ++ii) {
                Acc acc;
                acc.init(mat(i + ii, 0), v[0]);
#pragma unroll
                for (int j = 1; j < D; ++j) acc.mac(mat(i + ii, j), v[j]);
                qc[ii] = acc.result();
            }
            *reinterpret_cast<Unit16*>(tile + (base ^ tval[i])) = q;
        }
    } else {
#pragma unroll
        for (int i = 0; i < D; ++i) {
            Acc acc;
            acc.init(mat(i, 0), v[0]);
#pragma unroll
            for (int j = 1; j < D; ++j) acc.mac(mat(i, j), v[j]);
            tile[base ^ tval[i]] = acc.result();
        }
    }
}

// All groups of a tile owned by thread `tid` of `nthr` (power of two, tb = log2(nthr)).
template <typename C, int K, typename Mat>
TCB_HD void apply_block_on_tile(C* tile, const GroupMap& gm, int tid, int nthr, int tb,
                                const Mat& mat) {
    const uint32_t ngroups = 1u << gm.ngb;
    if ((uint32_t)tid >= ngroups) return;
    // bits of the group index that come from tid are fixed for this thread
    const uint32_t base_tid = group_base(gm, (uint32_t)tid);
    const bool vec0 = gm.vec0 != 0;
    for (uint32_t it = 0; ((it << tb) | (uint32_t)tid) < ngroups; ++it) {
        uint32_t b = base_tid;
        for (int i = tb; i < gm.ngb; ++i) b ^= (0u - ((it >> (i - tb)) & 1u)) & gm.ntval[i];
        apply_group<C, K>(tile, b, gm.tval, vec0, mat);
    }
}

// Fast path: 256 threads and exactly 2^NITLOG groups per thread (ngb == 8 + NITLOG).  The
// thread's part of the group index is folded once; the per-iteration part walks a Gray code, so
// each further group costs one XOR; everything is unrolled with compile-time table indices.
template <typename C, int K, int NITLOG, bool VEC0, typename Mat, int TB = 8, bool SPLIT = false>
TCB_HD void apply_block_fast_v(C* tile, const GroupMap& gm, int tid, const Mat& mat) {
    uint32_t b = 0;
#pragma unroll
    for (int i = 0; i < TB; ++i) b ^= (0u - (((uint32_t)tid >> i) & 1u)) & gm.ntval[i];
    uint32_t hv[NITLOG > 0 ? NITLOG : 1];
#pragma unroll
    for (int i = 0; i < NITLOG; ++i) hv[i] = gm.ntval[TB + i];
    uint32_t tv[1 << K];
#pragma unroll
    for (int j = 0; j < (1 << K); ++j) tv[j] = gm.tval[j];
#pragma unroll
    for (int it = 0; it < (1 << NITLOG); ++it) {
        if (it > 0) {
            int z = 0;
#pragma unroll
            for (int q = 0; q < NITLOG; ++q)
                if (((it >> q) & 1) && ((it & ((1 << q) - 1)) == 0)) z = q;  // count trailing zeros
            b ^= hv[z];
        }
        apply_group<C, K, Mat, SPLIT>(tile, b, tv, VEC0, mat);
    }
}

template <typename C, int K, int NITLOG, typename Mat, int TB = 8, bool SPLIT = false>
TCB_HD void apply_block_fast(C* tile, const GroupMap& gm, int tid, const Mat& mat) {
    if (sizeof(C) == 8 && gm.vec0)
        apply_block_fast_v<C, K, NITLOG, true, Mat, TB, SPLIT>(tile, gm, tid, mat);
    else
        apply_block_fast_v<C, K, NITLOG, false, Mat, TB, SPLIT>(tile, gm, tid, mat);
}

// Same for the persistent TMA pass (2^TB threads per CTA, tile size fixed at compile time):
// 2^(T-K-TB) groups per thread, generic loop when a wide block leaves less than one per thread.
template <typename C, int K, int T, int TB, typename Mat>
TCB_HD void apply_block_tb(C* tile, const GroupMap& gm, int tid, const Mat& mat) {
    if constexpr (T - K - TB >= 0) {
        if (gm.ngb == T - K) {
            apply_block_fast<C, K, T - K - TB, Mat, TB, true>(tile, gm, tid, mat);
            return;
        }
    }
    apply_block_on_tile<C, K>(tile, gm, tid, 1 << TB, TB, mat);
}

// Picks the fast path when the launch shape allows it (the production tiles: 2^12 / 2^13
// amplitudes, 256 threads), the generic loop otherwise (small states, test tile sizes).
template <typename C, int K, typename Mat>
TCB_HD void apply_block_dispatch(C* tile, const GroupMap& gm, int tid, int nthr, int tb, const Mat& mat) {
    if (nthr == 256) {
        const int nl = gm.ngb - 8;
        if (nl == 11 - K - 8 && 11 - K - 8 >= 0) {
            apply_block_fast<C, K, (11 - K - 8 >= 0 ? 11 - K - 8 : 0)>(tile, gm, tid, mat);
            return;
        }
        if (nl == 12 - K - 8 && 12 - K - 8 >= 0) {
            apply_block_fast<C, K, (12 - K - 8 >= 0 ? 12 - K - 8 : 0)>(tile, gm, tid, mat);
            return;
        }
        if (nl == 13 - K - 8) {
            apply_block_fast<C, K, 13 - K - 8>(tile, gm, tid, mat);
            return;
        }
    }
    apply_block_on_tile<C, K>(tile, gm, tid, nthr, tb, mat);
}

// ---- register tiles: several small gates per shared-memory round trip -----------------------
// A register tile is a set of KT <= 4 tile-local bits.  A thread loads the 2^KT amplitudes of a
// group once, applies a *sequence* of 1- and 2-bit gates that live inside those bits entirely
// in registers, and stores the group once.  Compared with one shared-memory round trip per
// gate this divides the LDS/STS traffic and the addressing overhead by the number of gates in
// the tile, which is what a staged pass is bound by once the FP32 pipe is busy.
struct RSub {
    int k;     // 1 or 2
    int p0;    // position (0..KT-1) of matrix index bit 0 inside the register tile
    int p1;    // position of matrix index bit 1 (k == 2), p1 > p0
    int moff;  // offset of the 2^k x 2^k matrix in the pass blob (complex elements, even)
};

struct RTile {
    int kt;    // register-tile bits (1..4)
    int nsub;  // gates applied inside the tile
    int sub0;  // index of the first RSub
    GroupMap gm;
};

// one 2-bit gate on positions (P0, P1) of the 2^KT register array
template <typename C, int KT, int P0, int P1>
TCB_HD void rsub2r(C* v, const C* mr);

template <typename C, int KT, int P0, int P1>
TCB_HD void rsub2(C* v, const C* m) {
    C mr[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) mr[i] = m[i];
    rsub2r<C, KT, P0, P1>(v, mr);
}

// same with the matrix already in registers
template <typename C, int KT, int P0, int P1>
TCB_HD void rsub2r(C* v, const C* mr) {
#pragma unroll
    for (int r = 0; r < (1 << (KT - 2)); ++r) {
        // spread r over the positions other than P0, P1
        int base = 0, rb = 0;
#pragma unroll
        for (int b = 0; b < KT; ++b)
            if (b != P0 && b != P1) {
                base |= ((r >> rb) & 1) << b;
                ++rb;
            }
        const int i0 = base, i1 = base | (1 << P0), i2 = base | (1 << P1), i3 = base | (1 << P0) | (1 << P1);
        const C a0 = v[i0], a1 = v[i1], a2 = v[i2], a3 = v[i3];
        RowAcc<C> o;
        o.init(mr[0], a0); o.mac(mr[1], a1); o.mac(mr[2], a2); o.mac(mr[3], a3); v[i0] = o.result();
        o.init(mr[4], a0); o.mac(mr[5], a1); o.mac(mr[6], a2); o.mac(mr[7], a3); v[i1] = o.result();
        o.init(mr[8], a0); o.mac(mr[9], a1); o.mac(mr[10], a2); o.mac(mr[11], a3); v[i2] = o.result();
        o.init(mr[12], a0); o.mac(mr[13], a1); o.mac(mr[14], a2); o.mac(mr[15], a3); v[i3] = o.result();
    }
}

// one 1-bit gate on position P0
template <typename C, int KT, int P0>
TCB_HD void rsub1(C* v, const C* m) {
    const C m0 = m[0], m1 = m[1], m2 = m[2], m3 = m[3];
#pragma unroll
    for (int r = 0; r < (1 << (KT - 1)); ++r) {
        const int lo = r & ((1 << P0) - 1);
        const int i0 = ((r >> P0) << (P0 + 1)) | lo, i1 = i0 | (1 << P0);
        const C a0 = v[i0], a1 = v[i1];
        RowAcc<C> o;
        o.init(m0, a0); o.mac(m1, a1); v[i0] = o.result();
        o.init(m2, a0); o.mac(m3, a1); v[i1] = o.result();
    }
}

template <typename C, int KT>
TCB_HD void rsub_dispatch(C* v, const RSub& s, const C* bm) {
    const C* m = bm + s.moff;
    if (s.k == 1) {
        switch (s.p0) {
            case 0: rsub1<C, KT, 0>(v, m); break;
            case 1: if constexpr (KT > 1) rsub1<C, KT, 1>(v, m); break;
            case 2: if constexpr (KT > 2) rsub1<C, KT, 2>(v, m); break;
            default: if constexpr (KT > 3) rsub1<C, KT, 3>(v, m); break;
        }
        return;
    }
    switch (s.p0 * 4 + s.p1) {
        case 1: if constexpr (KT > 1) rsub2<C, KT, 0, 1>(v, m); break;
        case 2: if constexpr (KT > 2) rsub2<C, KT, 0, 2>(v, m); break;
        case 3: if constexpr (KT > 3) rsub2<C, KT, 0, 3>(v, m); break;
        case 6: if constexpr (KT > 2) rsub2<C, KT, 1, 2>(v, m); break;
        case 7: if constexpr (KT > 3) rsub2<C, KT, 1, 3>(v, m); break;
        default: if constexpr (KT > 3) rsub2<C, KT, 2, 3>(v, m); break;  // 11
    }
}

// load one group, run the tile's gates in registers, store the group
template <typename C, int KT, bool VEC0>
TCB_HD void rtile_group(C* tile, uint32_t base, const uint32_t* tv, const RTile& rt, const RSub* subs, const C* bm) {
    constexpr int D = 1 << KT;
    C v[D];
    if (VEC0) {
#pragma unroll
        for (int j = 0; j < D; j += 2) {
            const Unit16 q = *reinterpret_cast<const Unit16*>(tile + (base ^ tv[j]));
            const C* qc = reinterpret_cast<const C*>(&q);
            v[j] = qc[0];
            v[(j + 1) % D] = qc[1 % (16 / (int)sizeof(C))];
        }
    } else {
#pragma unroll
        for (int j = 0; j < D; ++j) v[j] = tile[base ^ tv[j]];
    }
    for (int s = 0; s < rt.nsub; ++s) rsub_dispatch<C, KT>(v, subs[rt.sub0 + s], bm);
    if (VEC0) {
#pragma unroll
        for (int j = 0; j < D; j += 2) {
            Unit16 q;
            C* qc = reinterpret_cast<C*>(&q);
            qc[0] = v[j];
            qc[1 % (16 / (int)sizeof(C))] = v[(j + 1) % D];
            *reinterpret_cast<Unit16*>(tile + (base ^ tv[j])) = q;
        }
    } else {
#pragma unroll
        for (int j = 0; j < D; ++j) tile[base ^ tv[j]] = v[j];
    }
}

template <typename C, int KT, bool VEC0>
TCB_HD void rtile_run_v(C* tile, const RTile& rt, const RSub* subs, const C* bm, int tid, int nthr, int tb) {
    const GroupMap& gm = rt.gm;
    const uint32_t ngroups = 1u << gm.ngb;
    if ((uint32_t)tid >= ngroups) return;
    uint32_t tv[1 << KT];
#pragma unroll
    for (int j = 0; j < (1 << KT); ++j) tv[j] = gm.tval[j];
    uint32_t b = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) b ^= (0u - (((uint32_t)tid >> i) & 1u)) & gm.ntval[i];
    if (nthr == 256) {
        // remaining group bits in Gray-code order: one XOR per further group
        const uint32_t nit = ngroups >> 8;
        for (uint32_t it = 0; it < (nit ? nit : 1u); ++it) {
            if (it > 0) {
                int z = 0;
                while (!((it >> z) & 1u)) ++z;
                b ^= gm.ntval[8 + z];
            }
            rtile_group<C, KT, VEC0>(tile, b, tv, rt, subs, bm);
        }
    } else {
        for (uint32_t it = 0; ((it << tb) | (uint32_t)tid) < ngroups; ++it) {
            const uint32_t g = (it << tb) | (uint32_t)tid;
            uint32_t bb = 0;
            for (int i = 0; i < gm.ngb; ++i) bb ^= (0u - ((g >> i) & 1u)) & gm.ntval[i];
            rtile_group<C, KT, VEC0>(tile, bb, tv, rt, subs, bm);
        }
    }
}

// Pair path: a 4-bit register tile holding exactly two 2-bit gates on disjoint positions
// (P0,P1) and the complementary pair.  Both matrices are loaded once per thread, each group of
// 16 amplitudes makes one shared-memory round trip for 32 FMA per amplitude -- twice the
// arithmetic intensity of running the two gates as separate blocks.
template <typename C, int P0, int P1, bool VEC0>
TCB_HD void rpair_run(C* tile, const RTile& rt, const C* mA, const C* mB, int tid) {
    constexpr int Q0 = (P0 != 0 && P1 != 0) ? 0 : ((P0 != 1 && P1 != 1) ? 1 : 2);
    constexpr int Q1 = (P0 != 3 && P1 != 3) ? 3 : ((P0 != 2 && P1 != 2) ? 2 : 1);
    const GroupMap& gm = rt.gm;
    C ra[16], rb[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        ra[i] = mA[i];
        rb[i] = mB[i];
    }
    uint32_t tv[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) tv[j] = gm.tval[j];
    uint32_t b = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) b ^= (0u - (((uint32_t)tid >> i) & 1u)) & gm.ntval[i];
    const uint32_t nit = (1u << gm.ngb) >> 8;
    for (uint32_t it = 0; it < nit; ++it) {
        if (it > 0) {
            int z = 0;
            while (!((it >> z) & 1u)) ++z;
            b ^= gm.ntval[8 + z];
        }
        C v[16];
        if (VEC0) {
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                const Unit16 q = *reinterpret_cast<const Unit16*>(tile + (b ^ tv[j]));
                const C* qc = reinterpret_cast<const C*>(&q);
                v[j] = qc[0];
                v[j + 1] = qc[1 % (16 / (int)sizeof(C))];
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = tile[b ^ tv[j]];
        }
        rsub2r<C, 4, P0, P1>(v, ra);
        rsub2r<C, 4, Q0, Q1>(v, rb);
        if (VEC0) {
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                Unit16 q;
                C* qc = reinterpret_cast<C*>(&q);
                qc[0] = v[j];
                qc[1 % (16 / (int)sizeof(C))] = v[j + 1];
                *reinterpret_cast<Unit16*>(tile + (b ^ tv[j])) = q;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) tile[b ^ tv[j]] = v[j];
        }
    }
}

template <typename C, int P0, int P1>
TCB_HD void rpair_dispatch(C* tile, const RTile& rt, const C* mA, const C* mB, int tid) {
    if (sizeof(C) == 8 && rt.gm.vec0)
        rpair_run<C, P0, P1, true>(tile, rt, mA, mB, tid);
    else
        rpair_run<C, P0, P1, false>(tile, rt, mA, mB, tid);
}

template <typename C>
TCB_HD void rtile_run(C* tile, const RTile& rt, const RSub* subs, const C* bm, int tid, int nthr, int tb) {
    const bool vec0 = sizeof(C) == 8 && rt.gm.vec0 != 0;
    const RSub s0 = subs[rt.sub0];
    if (rt.nsub == 1 && rt.kt == s0.k) {
        // one gate covering the whole tile: the plain block path (matrix in registers)
        const C* m = bm + s0.moff;
        if (s0.k == 1) {
            C r[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) r[i] = m[i];
            apply_block_dispatch<C, 1>(tile, rt.gm, tid, nthr, tb, [&](int i, int j) { return r[i * 2 + j]; });
        } else {
            C r[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = m[i];
            apply_block_dispatch<C, 2>(tile, rt.gm, tid, nthr, tb, [&](int i, int j) { return r[i * 4 + j]; });
        }
        return;
    }
    if constexpr (sizeof(C) == 8) {
        if (rt.kt == 4 && rt.nsub == 2 && nthr == 256 && rt.gm.ngb >= 8) {
            const RSub s1 = subs[rt.sub0 + 1];
            const int ma = (1 << s0.p0) | (1 << s0.p1), mb = (1 << s1.p0) | (1 << s1.p1);
            if (s0.k == 2 && s1.k == 2 && (ma & mb) == 0) {
                const C* mA = bm + s0.moff;
                const C* mB = bm + s1.moff;
                switch (s0.p0 * 4 + s0.p1) {
                    case 1: rpair_dispatch<C, 0, 1>(tile, rt, mA, mB, tid); break;
                    case 2: rpair_dispatch<C, 0, 2>(tile, rt, mA, mB, tid); break;
                    case 3: rpair_dispatch<C, 0, 3>(tile, rt, mA, mB, tid); break;
                    case 6: rpair_dispatch<C, 1, 2>(tile, rt, mA, mB, tid); break;
                    case 7: rpair_dispatch<C, 1, 3>(tile, rt, mA, mB, tid); break;
                    default: rpair_dispatch<C, 2, 3>(tile, rt, mA, mB, tid); break;
                }
                return;
            }
        }
    }
    switch (rt.kt) {
        case 1:
            rtile_run_v<C, 1, false>(tile, rt, subs, bm, tid, nthr, tb);
            break;
        case 2:
            if (vec0) rtile_run_v<C, 2, true>(tile, rt, subs, bm, tid, nthr, tb);
            else rtile_run_v<C, 2, false>(tile, rt, subs, bm, tid, nthr, tb);
            break;
        case 3:
            if (vec0) rtile_run_v<C, 3, true>(tile, rt, subs, bm, tid, nthr, tb);
            else rtile_run_v<C, 3, false>(tile, rt, subs, bm, tid, nthr, tb);
            break;
        default:
            // 4-bit register tiles only for complex64 (16 amplitudes = 32 registers); for
            // complex128 the host plans <= 3 bits and abi checks it
            if constexpr (sizeof(C) == 8) {
                if (vec0) rtile_run_v<C, 4, true>(tile, rt, subs, bm, tid, nthr, tb);
                else rtile_run_v<C, 4, false>(tile, rt, subs, bm, tid, nthr, tb);
            }
            break;
    }
}

// ---- 16-amplitude register tiles of the persistent pass (trpass_kernel, complex64) ------------
// Every op of a trpass is a 4-bit register tile: a thread owns ONE group of 16 amplitudes, loads
// it once, applies the op's 1- and 2-bit gates on positions 0..3 of the register array, and
// stores it once.  A plain 2-bit block is the special case "one gate, two filler positions" --
// the fillers are bits a thread would otherwise iterate over.
//
// Control is warp-uniform and lives in the kernel-parameter constant bank: per op a TROp, per
// gate one code word  (case | matrix offset << 8),  case 0..5 = 2-bit gate on positions
// (0,1) (0,2) (0,3) (1,2) (1,3) (2,3), case 6..9 = 1-bit gate on position 0..3.  Every gate owns a
// 16-element matrix slot, so the matrix of the first gate is fetched together with the
// amplitudes, before its code is looked at (no load -> branch -> load chain at the op start).
// Per-op control word block (16 bytes, kernel-parameter constant bank).  The compute loop fetches
// the block of op o+1 while op o is in the FP32 pipe, so an op starts with everything but its
// amplitudes and its first matrix already in registers.
struct TROp {
    uint32_t t01;    // swizzled BYTE offsets of positions 0 and 1 inside the tile (16 bits each)
    uint32_t t23;    // ... of positions 2 and 3
    uint32_t flags;  // nsub | vec0 << 8 | sync << 12 | sub0 << 16
    uint32_t code0;  // code word of the first gate
};
// vec0: position 0 is amplitude bit 0, (j, j+1) share an aligned 16-byte unit
// sync (before this op): 2 = CTA barrier (new segment), 1 = __syncwarp() (same segment)
TCB_HD int trop_nsub(const TROp& op) { return (int)(op.flags & 0xffu); }
TCB_HD int trop_vec0(const TROp& op) { return (int)((op.flags >> 8) & 1u); }
TCB_HD int trop_sync(const TROp& op) { return (int)((op.flags >> 12) & 3u); }
TCB_HD int trop_sub0(const TROp& op) { return (int)(op.flags >> 16); }
TCB_HD uint32_t trop_t8(const TROp& op, int i) {
    const uint32_t w = i < 2 ? op.t01 : op.t23;
    return (i & 1) ? (w >> 16) : (w & 0xffffu);
}

TCB_HD void tr_load_matrix(float2* mr, const float2* slot) {
    const float4* m4 = reinterpret_cast<const float4*>(slot);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 q = m4[i];
        mr[2 * i] = make_float2(q.x, q.y);
        mr[2 * i + 1] = make_float2(q.z, q.w);
    }
}

// one gate (case word c, common.cuh above) on the 16 amplitudes in registers; an if-tree, not a
// switch: a jump table would put one more dependent constant load in front of the first FFMA2
TCB_HD void tr_gate(float2* v, const float2* mr, uint32_t c) {
    if (c < 3u) {
        if (c == 0u) rsub2r<float2, 4, 0, 1>(v, mr);
        else if (c == 1u) rsub2r<float2, 4, 0, 2>(v, mr);
        else rsub2r<float2, 4, 0, 3>(v, mr);
    } else if (c < 6u) {
        if (c == 3u) rsub2r<float2, 4, 1, 2>(v, mr);
        else if (c == 4u) rsub2r<float2, 4, 1, 3>(v, mr);
        else rsub2r<float2, 4, 2, 3>(v, mr);
    } else if (c < 8u) {
        if (c == 6u) rsub1<float2, 4, 0>(v, mr);
        else rsub1<float2, 4, 1>(v, mr);
    } else {
        if (c == 8u) rsub1<float2, 4, 2>(v, mr);
        else rsub1<float2, 4, 3>(v, mr);
    }
}

// tile: shared-memory tile (bytes); b8: swizzled byte offset of the thread's group (element 0)
TCB_HD void trtile_thread(unsigned char* tile, uint32_t b8, const TROp op, const uint32_t* subcode, const float2* bm) {
    uint32_t a[16];
    a[0] = b8;
#pragma unroll
    for (int j = 1; j < 16; ++j) {
        const int low = j & (-j);  // lowest set bit: a[j] = a[j - low] ^ t8[log2(low)]
        a[j] = a[j ^ low] ^ trop_t8(op, low == 1 ? 0 : (low == 2 ? 1 : (low == 4 ? 2 : 3)));
    }
    const bool vec0 = trop_vec0(op) != 0;
    float2 v[16];
    if (vec0) {
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
            const float4 q = *reinterpret_cast<const float4*>(tile + a[j]);
            v[j] = make_float2(q.x, q.y);
            v[j + 1] = make_float2(q.z, q.w);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = *reinterpret_cast<const float2*>(tile + a[j]);
    }
    uint32_t code = op.code0;
    float2 mr[16];
    tr_load_matrix(mr, bm + (code >> 8));
    const int nsub = trop_nsub(op), sub0 = trop_sub0(op);
    for (int s = 0;;) {
        const uint32_t next = (s + 1 < nsub) ? subcode[sub0 + s + 1] : 0u;  // in flight during the FMAs
        tr_gate(v, mr, code & 15u);
        if (++s >= nsub) break;
        code = next;
        tr_load_matrix(mr, bm + (code >> 8));
    }
    if (vec0) {
#pragma unroll
        for (int j = 0; j < 16; j += 2)
            *reinterpret_cast<float4*>(tile + a[j]) = make_float4(v[j].x, v[j].y, v[j + 1].x, v[j + 1].y);
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) *reinterpret_cast<float2*>(tile + a[j]) = v[j];
    }
}


// ---- TMA decomposition of a tile (tpass_kernel) ----------------------------------------------
// The state is described to the TMA unit as a rank-5 tensor of 128-byte lines:
//   d0 = the scalars of one line, d1 = line index over the whole buffer (stride 128 B),
//   d2..d4 = up to three "stride bits" (size 2, stride 2^bit amplitudes).
// One box = 2^lrow2 contiguous amplitudes x the box's stride bits; the remaining stride bits of
// the tile are enumerated as 2^nextra boxes.  Stride bits are the gathered bits of the tile,
// preceded by the top bits of a row that is longer than 256 lines.  Boxes land back to back in
// shared memory, so the local amplitude index is exactly the TileGeom one, stored in the
// SWZ_HW128 layout.
struct TmaPlan {
    int line_bits;      // log2(amplitudes per 128-byte line): 4 (complex64) / 3 (complex128)
    int lrow2;          // contiguous low bits of one box row (<= line_bits + 8)
    int nbox_bits;      // stride bits inside one box (<= 3)
    int box_bit[3];     // their global bit positions
    int nextra;         // stride bits enumerated across boxes
    int extra_bit[8];   // their global bit positions
    int box_amps_log2;  // lrow2 + nbox_bits
};

// 0 when the tile can be moved by TMA boxes, -1 otherwise (rows shorter than 8 lines)
inline int make_tma_plan(const TileGeom& g, int apu, TmaPlan* tp) {
    tp->line_bits = (apu == 2) ? 4 : 3;
    if (g.lrow < tp->line_bits) return -1;  // a row is at least one 128-byte line
    tp->lrow2 = g.lrow < tp->line_bits + 8 ? g.lrow : tp->line_bits + 8;
    int sb[16];
    int ns = 0;
    for (int b = tp->lrow2; b < g.lrow; ++b) sb[ns++] = b;
    for (int j = 0; j < g.h; ++j) sb[ns++] = g.hb[j];
    tp->nbox_bits = ns < 3 ? ns : 3;
    for (int i = 0; i < 3; ++i) tp->box_bit[i] = i < tp->nbox_bits ? sb[i] : -1;
    tp->nextra = ns - tp->nbox_bits;
    if (tp->nextra > 8) return -1;
    for (int i = 0; i < 8; ++i) tp->extra_bit[i] = i < tp->nextra ? sb[tp->nbox_bits + i] : -1;
    tp->box_amps_log2 = tp->lrow2 + tp->nbox_bits;
    return 0;
}

// amplitude offset (relative to the tile base) of box c
TCB_HD uint64_t tma_box_offset(const TmaPlan& tp, uint32_t c) {
    uint64_t o = 0;
    for (int i = 0; i < tp.nextra; ++i)
        if ((c >> i) & 1u) o |= 1ull << tp.extra_bit[i];
    return o;
}

// ---- host helpers (plan.cpp part of abi.cu) ------------------------------------------------
// Choose the tile for a set of ascending target bits: gathers exactly the targets that do not
// fall into the contiguous low part.  Returns <0 on error.
int make_geom(int nbits, int tile_bits, int k, const int* bits, TileGeom* g);
// Tile with explicitly requested gathered bits.
int make_geom_hi(int nbits, int tile_bits, int n_hi, const int* tile_hi, TileGeom* g, int max_hi = 8);
// Fill the GroupMap for a block with the given ascending global bits inside geometry g.
int make_group_map(const TileGeom& g, int apu, int k, const int* bits, GroupMap* gm, int swz_mode = SWZ_SW);
// Persistent TMA pipeline for a multi-block pass (tpass.cu).  0: launched; > 0: not eligible
// (tile size overridden, state smaller than a tile, no driver entry point, TCB200_TMA=0) and the
// caller falls back to cpass_kernel; < 0: error.
int launch_tpass(int dtype, void* state, int nbits, int nops, const int* ops_k, const int* ops_bits,
                 const double* mats, int n_hi, const int* tile_hi, int64_t batch, cudaStream_t st);
// tensor map (CUtensorMap, 128 bytes, 64-byte aligned) + box plan for tiles of geometry g; 0: ready, > 0: not eligible
int tma_encode_state_map(void* state, int nbits, int is_c64, const TileGeom& g, int64_t batch, TmaPlan* tp, void* map_out);
// complex64 register-tile pass (<= 4-bit tiles of 1-/2-bit gates) through the same pipeline
int launch_trpass(void* state, int nbits, int nrt, const int* rt_k, const int* rt_bits, const int* rt_nsub,
                  const int* sub_k, const int* sub_bits, const double* sub_mats, int n_hi, const int* tile_hi,
                  int64_t batch, cudaStream_t st);
int pick_threads(int T, int k, int apu);   // log2(blockDim.x)
int dense_tile_bits(int dtype, int k);     // log2(tile amplitudes) of the single-block kernel
int pass_tile_bits(int dtype);             // ... of the multi-block pass kernel
int expect_tile_bits(int dtype);           // ... of the expectation kernel

// ---- expectation: one term on one staged tile, elements owned by thread tid ------------------
template <typename C, typename R>
TCB_HD void expect_tile_term(const C* tile, uint32_t tsz, uint32_t fl, uint32_t sl, int tid, int nthr,
                             R* out_re, R* out_im) {
    R pr = 0, pi = 0;
    for (uint32_t e = tid; e < tsz; e += nthr) {
        const C a = tile[e];
        const C b = tile[e ^ fl];
        R re = a.x * b.x + a.y * b.y;
        R im = a.x * b.y - a.y * b.x;
        uint32_t par = e & sl;
        par ^= par >> 16;
        par ^= par >> 8;
        par ^= par >> 4;
        par ^= par >> 2;
        par ^= par >> 1;
        if (par & 1u) {
            re = -re;
            im = -im;
        }
        pr += re;
        pi += im;
    }
    *out_re = pr;
    *out_im = pi;
}

TCB_HD uint32_t parity32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__popc(x) & 1u;
#else
    x ^= x >> 16;
    x ^= x >> 8;
    x ^= x >> 4;
    x ^= x >> 2;
    x ^= x >> 1;
    return x & 1u;
#endif
}

// One term on the NI amplitudes a thread holds in registers (elements e0 + i * nthr).
// ODD: the imaginary component is the non-zero one; SIGNED: the local sign mask is not empty;
// GUARD: the tile may end inside the NI elements (small states only).
template <typename C, typename R, int NI, bool ODD, bool SIGNED, bool GUARD>
TCB_HD R expect_term_loop(const C* a, const C* tile, uint32_t tsz, uint32_t e0, uint32_t nthr, uint32_t f, uint32_t s) {
    R acc = 0;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        const uint32_t e = e0 + (uint32_t)i * nthr;
        if (GUARD && e >= tsz) break;
        const C b = tile[e ^ f];
        R d;
        if (ODD) {
            d = a[i].x * b.y;
            d = fma(-a[i].y, b.x, d);
        } else {
            d = a[i].x * b.x;
            d = fma(a[i].y, b.y, d);
        }
        // popc(e & s) = popc(e0 & s) + popc((i * nthr) & s) mod 2: nthr is a power of two above tid
        if (SIGNED && parity32(((uint32_t)i * nthr) & s)) d = -d;
        acc += d;
    }
    if (SIGNED && parity32(e0 & s)) acc = -acc;
    return acc;
}

// All terms of a launch on one staged tile.  The thread's own amplitudes (element tid + i * nthr)
// are loaded once into registers and reused by every term; a term then costs one partner load
// and two FMAs per amplitude: S = sum_e conj(psi_e) psi_{e^f} (-1)^{popc(e & s)} is real when the
// string holds an even number of Y's and imaginary when odd (the string is Hermitian and the
// phase (-i)^{n_y} is applied afterwards), so only that component is formed (pv[t]) -- bit t of
// `odd_mask` selects the imaginary one -- and the other is exactly 0.
template <typename C, typename R, int MT>
TCB_HD void expect_tile_terms(const C* tile, uint32_t tsz, int nterms, const uint32_t* fl, const uint32_t* sl,
                              uint32_t odd_mask, int tid, int nthr, R* pv) {
    constexpr int NI = 8;  // amplitudes held in registers at a time (the 32 KiB tile gives a thread 16 / 8)
#pragma unroll
    for (int t = 0; t < MT; ++t) pv[t] = 0;
    const bool whole = (tsz % (NI * (uint32_t)nthr)) == 0;
    for (uint32_t e0 = (uint32_t)tid; e0 < tsz; e0 += NI * (uint32_t)nthr) {
        C a[NI];
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const uint32_t e = e0 + (uint32_t)i * (uint32_t)nthr;
            a[i] = (whole || e < tsz) ? tile[e] : mk<C, R>(0, 0);
        }
#pragma unroll
        for (int t = 0; t < MT; ++t) {
            if (t < nterms) {
                const uint32_t f = fl[t], s = sl[t], nt = (uint32_t)nthr;
                const bool odd = (odd_mask >> t) & 1u;
                R acc;
                if (!whole) {
                    acc = odd ? expect_term_loop<C, R, NI, true, true, true>(a, tile, tsz, e0, nt, f, s)
                              : expect_term_loop<C, R, NI, false, true, true>(a, tile, tsz, e0, nt, f, s);
                } else if (s == 0) {
                    acc = odd ? expect_term_loop<C, R, NI, true, false, false>(a, tile, tsz, e0, nt, f, s)
                              : expect_term_loop<C, R, NI, false, false, false>(a, tile, tsz, e0, nt, f, s);
                } else {
                    acc = odd ? expect_term_loop<C, R, NI, true, true, false>(a, tile, tsz, e0, nt, f, s)
                              : expect_term_loop<C, R, NI, false, true, false>(a, tile, tsz, e0, nt, f, s);
                }
                pv[t] += acc;
            }
        }
    }
}

TCB_HD bool parity64(uint64_t x) {
    x ^= x >> 32;
    x ^= x >> 16;
    x ^= x >> 8;
    x ^= x >> 4;
    x ^= x >> 2;
    x ^= x >> 1;
    return (x & 1ull) != 0;
}

}  // namespace tcb
