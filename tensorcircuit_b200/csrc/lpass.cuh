// Structure-aware staged pass ("gate pass"): device bodies shared by lpass_kernel (lpass.cu) and
// the CPU emulation in tests/emu (everything that computes an address is __host__ __device__).
//
// A pass keeps one tile of 2^T amplitudes in shared memory and runs a list of ROUNDS on it.  In a
// round every thread owns groups of 16 amplitudes -- 4 "register bits" of the tile-local index --
// loads a group once, applies the round's micro-ops to it entirely in registers, and stores it
// once: one shared-memory round trip for up to LP_MAX_CODES gates instead of one per gate.
//
// The tile is addressed through a GF(2)-affine index map maintained by the host,
//     byte offset of logical tile index x  =  d  ^  XOR_t  x_t * col[t],
// which starts as the staging swizzle (common.cuh: swz_unit) and absorbs every gate whose matrix
// is a permutation matrix of an affine bit map (cnot, swap, x, cx ladders ...): such a gate only
// rewrites `col` / `d` between two rounds and costs no instruction on the device.  The final
// map is undone by the write-back (lstage_out_thread), so a pass is exact and self-contained.
//
// Micro-ops of a round act on register positions 0..3 (the round's register bits in ascending
// logical order): general 2x2 on one position, 4x4 on a pair, 8x8 on three positions, a 16-entry
// diagonal table.  Matrices sit in the kernel-parameter constant bank and reach the FMAs as
// uniform-register operands (LDCU -> UR): complex64 elements are stored as (re, -, im, im) so that
// both FFMA2 of a complex multiply-add take their matrix operand exactly as loaded -- a broadcast
// scalar for (re, re) * (x, y) and a 64-bit pair for (-im, im) * (y, x) with the swap and the sign
// folded into operand modifiers.  No MOV is spent on the matrix and no vector register holds it.
#pragma once

#include "common.cuh"

namespace tcb {

constexpr int LP_RB = 4;           // register bits per round (16 amplitudes per group)
constexpr int LP_MAX_ROUNDS = 48;  // rounds per pass
constexpr int LP_MAX_CODES = 12;   // micro-ops per round
constexpr int LP_MAX_GB = 9;       // group-index bits: T - LP_RB for the 64 KiB complex64 tile
constexpr int LP_MAX_T = LP_MAX_GB + LP_RB;
constexpr int LP_MAT_ELEMS = 1280;  // matrix elements (16 B each) per pass: 80 4x4 blocks

// opcodes (low 8 bits of a code word; the rest is the offset of the op's matrix in ME units)
constexpr uint32_t LOP_G1 = 0;    // +p          general 2x2 on register position p (0..3)
constexpr uint32_t LOP_G2 = 4;    // +pair index general 4x4 on positions (0,1)(0,2)(0,3)(1,2)(1,3)(2,3)
constexpr uint32_t LOP_G3 = 10;   // +e          general 8x8 on the three positions other than e
constexpr uint32_t LOP_DG = 14;   //             16-entry diagonal table over the register index
constexpr uint32_t LOP_RA = 15;   // +p          2x2 with real diagonal and imaginary off-diagonal (rx, ...): half the FMAs
constexpr uint32_t LOP_RB = 19;   // +p          real 2x2 (h, ry, ...): half the FMAs
constexpr uint32_t LOP_COUNT = 23;

// one matrix element in the parameter bank
template <typename Real>
struct ME;
template <>
struct alignas(16) ME<float> {
    float re, pad, im0, im1;  // pad == re, im0 == im1: the (re, re) / (im, im) operand pairs of the two FFMA2
};
template <>
struct alignas(16) ME<double> {
    double re, im0;
};

struct LRound {
    uint32_t d;                  // byte offset of group 0, element 0
    uint32_t gcol[LP_MAX_GB];    // byte-offset masks of the group-index bits (host-chosen lane order)
    uint32_t rcol[LP_RB];        // byte-offset masks of the register positions
    uint32_t ncodes;             // bits 0..7 number of micro-ops, bit 8: 16-byte accesses (complex64 pairs)
    uint32_t code[LP_MAX_CODES];
};

struct LOut {                    // index map at the end of the pass, for the write-back
    uint32_t d;
    uint32_t col[LP_MAX_T];      // byte-offset mask of logical tile bit t
    uint32_t vec;                // complex64: (x, x^1) share one aligned 16-byte unit
};

// Staging constants of the production shape (256 threads, 16 units of 16 bytes per thread: the
// 64 KiB tile).  Thread `tid` owns the units tid + 256 i; everything that depends on i alone is
// precomputed on the host, so that a unit costs one add, one XOR and the copy itself.
constexpr int LP_FAST_ITERS = 16;
struct LStage {
    uint64_t goff[LP_FAST_ITERS];   // amplitude offset of unit (i << 8) relative to the tile base
    uint64_t bit_off[9];            // amplitude offset of bit b of the thread's element index tid * APU
    uint32_t sin[LP_FAST_ITERS];    // stage-in: byte offset of unit (i << 8) in the staging swizzle
    uint32_t sout[LP_FAST_ITERS];   // write-back: byte-offset XOR of unit (i << 8) under the final index map
};

// ---- complex arithmetic against a parameter-bank element ---------------------------------------
template <typename C, typename Real>
TCB_HD C me_mul(const ME<Real>& m, const C v) {
    C r;
    r.x = m.re * v.x - m.im0 * v.y;
    r.y = m.re * v.y + m.im0 * v.x;
    return r;
}
template <typename C, typename Real>
TCB_HD void me_fma(C& acc, const ME<Real>& m, const C v) {
    acc.x = fma(m.re, v.x, acc.x);
    acc.x = fma(-m.im0, v.y, acc.x);
    acc.y = fma(m.re, v.y, acc.y);
    acc.y = fma(m.im0, v.x, acc.y);
}
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
template <>
__device__ __forceinline__ float2 me_mul<float2, float>(const ME<float>& m, const float2 v) {
    const float2 acc = __fmul2_rn(make_float2(m.re, m.re), v);
    return __ffma2_rn(make_float2(-m.im0, m.im1), make_float2(v.y, v.x), acc);
}
template <>
__device__ __forceinline__ void me_fma<float2, float>(float2& acc, const ME<float>& m, const float2 v) {
    acc = __ffma2_rn(make_float2(m.re, m.re), v, acc);
    acc = __ffma2_rn(make_float2(-m.im0, m.im1), make_float2(v.y, v.x), acc);
}
#endif

// products with a purely real / purely imaginary element: one packed FMA instead of two
template <typename C, typename Real>
TCB_HD C me_mul_re(const ME<Real>& m, const C v) {
    C r;
    r.x = m.re * v.x;
    r.y = m.re * v.y;
    return r;
}
template <typename C, typename Real>
TCB_HD void me_fma_re(C& acc, const ME<Real>& m, const C v) {
    acc.x = fma(m.re, v.x, acc.x);
    acc.y = fma(m.re, v.y, acc.y);
}
template <typename C, typename Real>
TCB_HD C me_mul_im(const ME<Real>& m, const C v) {  // (i im) * v
    C r;
    r.x = -m.im0 * v.y;
    r.y = m.im0 * v.x;
    return r;
}
template <typename C, typename Real>
TCB_HD void me_fma_im(C& acc, const ME<Real>& m, const C v) {
    acc.x = fma(-m.im0, v.y, acc.x);
    acc.y = fma(m.im0, v.x, acc.y);
}
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
template <>
__device__ __forceinline__ float2 me_mul_re<float2, float>(const ME<float>& m, const float2 v) {
    return __fmul2_rn(make_float2(m.re, m.re), v);
}
template <>
__device__ __forceinline__ void me_fma_re<float2, float>(float2& acc, const ME<float>& m, const float2 v) {
    acc = __ffma2_rn(make_float2(m.re, m.re), v, acc);
}
template <>
__device__ __forceinline__ float2 me_mul_im<float2, float>(const ME<float>& m, const float2 v) {
    return __fmul2_rn(make_float2(-m.im0, m.im1), make_float2(v.y, v.x));
}
template <>
__device__ __forceinline__ void me_fma_im<float2, float>(float2& acc, const ME<float>& m, const float2 v) {
    acc = __ffma2_rn(make_float2(-m.im0, m.im1), make_float2(v.y, v.x), acc);
}
#endif

// ---- micro-ops on the 16 amplitudes of a group -------------------------------------------------
template <typename C, typename Real, int P0>
TCB_HD void lp_g1(C* v, const ME<Real>* m) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int lo = r & ((1 << P0) - 1);
        const int i0 = ((r >> P0) << (P0 + 1)) | lo, i1 = i0 | (1 << P0);
        const C a0 = v[i0], a1 = v[i1];
        C o0 = me_mul<C, Real>(m[0], a0);
        me_fma<C, Real>(o0, m[1], a1);
        C o1 = me_mul<C, Real>(m[2], a0);
        me_fma<C, Real>(o1, m[3], a1);
        v[i0] = o0;
        v[i1] = o1;
    }
}

// 2x2 whose elements are each purely real or purely imaginary.  IMOFF: the off-diagonal pair is
// imaginary (rx-like: cos I - i sin X); otherwise the whole matrix is real (h, ry).  4 real FMA per
// amplitude instead of 8.
template <typename C, typename Real, int P0, bool IMOFF>
TCB_HD void lp_r1(C* v, const ME<Real>* m) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int lo = r & ((1 << P0) - 1);
        const int i0 = ((r >> P0) << (P0 + 1)) | lo, i1 = i0 | (1 << P0);
        const C a0 = v[i0], a1 = v[i1];
        C o0 = me_mul_re<C, Real>(m[0], a0);
        C o1;
        if (IMOFF) {
            me_fma_im<C, Real>(o0, m[1], a1);
            o1 = me_mul_im<C, Real>(m[2], a0);
        } else {
            me_fma_re<C, Real>(o0, m[1], a1);
            o1 = me_mul_re<C, Real>(m[2], a0);
        }
        me_fma_re<C, Real>(o1, m[3], a1);
        v[i0] = o0;
        v[i1] = o1;
    }
}

template <typename C, typename Real, int P0, int P1>
TCB_HD void lp_g2(C* v, const ME<Real>* m) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        int base = 0, rb = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b)
            if (b != P0 && b != P1) {
                base |= ((r >> rb) & 1) << b;
                ++rb;
            }
        const int i0 = base, i1 = base | (1 << P0), i2 = base | (1 << P1), i3 = base | (1 << P0) | (1 << P1);
        const C a0 = v[i0], a1 = v[i1], a2 = v[i2], a3 = v[i3];
        C o;
        o = me_mul<C, Real>(m[0], a0); me_fma<C, Real>(o, m[1], a1); me_fma<C, Real>(o, m[2], a2); me_fma<C, Real>(o, m[3], a3); v[i0] = o;
        o = me_mul<C, Real>(m[4], a0); me_fma<C, Real>(o, m[5], a1); me_fma<C, Real>(o, m[6], a2); me_fma<C, Real>(o, m[7], a3); v[i1] = o;
        o = me_mul<C, Real>(m[8], a0); me_fma<C, Real>(o, m[9], a1); me_fma<C, Real>(o, m[10], a2); me_fma<C, Real>(o, m[11], a3); v[i2] = o;
        o = me_mul<C, Real>(m[12], a0); me_fma<C, Real>(o, m[13], a1); me_fma<C, Real>(o, m[14], a2); me_fma<C, Real>(o, m[15], a3); v[i3] = o;
    }
}

// 8x8 on the three positions other than E; matrix index bit j <-> the j-th of them (ascending)
template <typename C, typename Real, int E>
TCB_HD void lp_g3(C* v, const ME<Real>* m) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        int idx[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int e = s << E, jb = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if (b != E) {
                    e |= ((j >> jb) & 1) << b;
                    ++jb;
                }
            idx[j] = e;
        }
        C a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = v[idx[j]];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            C o = me_mul<C, Real>(m[i * 8], a[0]);
#pragma unroll
            for (int j = 1; j < 8; ++j) me_fma<C, Real>(o, m[i * 8 + j], a[j]);
            v[idx[i]] = o;
        }
    }
}

template <typename C, typename Real>
TCB_HD void lp_dg(C* v, const ME<Real>* m) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = me_mul<C, Real>(m[j], v[j]);
}

template <typename C, typename Real>
TCB_HD void lp_dispatch(C* v, uint32_t code, const ME<Real>* mats) {
    const ME<Real>* m = mats + (code >> 8);
    switch (code & 0xffu) {
        case 0: lp_g1<C, Real, 0>(v, m); break;
        case 1: lp_g1<C, Real, 1>(v, m); break;
        case 2: lp_g1<C, Real, 2>(v, m); break;
        case 3: lp_g1<C, Real, 3>(v, m); break;
        case 4: lp_g2<C, Real, 0, 1>(v, m); break;
        case 5: lp_g2<C, Real, 0, 2>(v, m); break;
        case 6: lp_g2<C, Real, 0, 3>(v, m); break;
        case 7: lp_g2<C, Real, 1, 2>(v, m); break;
        case 8: lp_g2<C, Real, 1, 3>(v, m); break;
        case 9: lp_g2<C, Real, 2, 3>(v, m); break;
        case 10: lp_g3<C, Real, 0>(v, m); break;
        case 11: lp_g3<C, Real, 1>(v, m); break;
        case 12: lp_g3<C, Real, 2>(v, m); break;
        case 13: lp_g3<C, Real, 3>(v, m); break;
        case 15: lp_r1<C, Real, 0, true>(v, m); break;
        case 16: lp_r1<C, Real, 1, true>(v, m); break;
        case 17: lp_r1<C, Real, 2, true>(v, m); break;
        case 18: lp_r1<C, Real, 3, true>(v, m); break;
        case 19: lp_r1<C, Real, 0, false>(v, m); break;
        case 20: lp_r1<C, Real, 1, false>(v, m); break;
        case 21: lp_r1<C, Real, 2, false>(v, m); break;
        case 22: lp_r1<C, Real, 3, false>(v, m); break;
        default: lp_dg<C, Real>(v, m); break;
    }
}

// ---- one group: load 16, run the round's micro-ops, store 16 -----------------------------------
// `vec` (complex64, rcol[0] == 8): (j, j+1) is one aligned 16-byte unit.  It is a run-time flag
// on purpose: the micro-op dispatch -- by far the largest piece of code -- must exist once in the
// kernel, or the instruction cache thrashes (four copies were 230 KB of SASS).
template <typename C, typename Real>
TCB_HD void lround_group(unsigned char* tile, uint32_t b, const LRound& R, const ME<Real>* mats, const bool vec) {
    const uint32_t c0 = R.rcol[0], c1 = R.rcol[1], c2 = R.rcol[2], c3 = R.rcol[3];
    const uint32_t nc = R.ncodes & 0xffu;
    C v[16];
    if (sizeof(C) == 8 && vec) {
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
            const uint32_t off = ((j & 2) ? c1 : 0u) ^ ((j & 4) ? c2 : 0u) ^ ((j & 8) ? c3 : 0u);
            const Unit16 q = *reinterpret_cast<const Unit16*>(tile + (b ^ off));
            const C* qc = reinterpret_cast<const C*>(&q);
            v[j] = qc[0];
            v[j + 1] = qc[1 % (16 / (int)sizeof(C))];
        }
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const uint32_t off = ((j & 1) ? c0 : 0u) ^ ((j & 2) ? c1 : 0u) ^ ((j & 4) ? c2 : 0u) ^ ((j & 8) ? c3 : 0u);
            v[j] = *reinterpret_cast<const C*>(tile + (b ^ off));
        }
    }
#pragma unroll 1
    for (uint32_t o = 0; o < nc; ++o) lp_dispatch<C, Real>(v, R.code[o], mats);
    if (sizeof(C) == 8 && vec) {
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
            const uint32_t off = ((j & 2) ? c1 : 0u) ^ ((j & 4) ? c2 : 0u) ^ ((j & 8) ? c3 : 0u);
            Unit16 q;
            C* qc = reinterpret_cast<C*>(&q);
            qc[0] = v[j];
            qc[1 % (16 / (int)sizeof(C))] = v[j + 1];
            *reinterpret_cast<Unit16*>(tile + (b ^ off)) = q;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const uint32_t off = ((j & 1) ? c0 : 0u) ^ ((j & 2) ? c1 : 0u) ^ ((j & 4) ? c2 : 0u) ^ ((j & 8) ? c3 : 0u);
            *reinterpret_cast<C*>(tile + (b ^ off)) = v[j];
        }
    }
}

// All groups of thread `tid` (2^tb threads, 2^ngb groups, ngb >= tb): the thread's bits of the
// group index are folded once, the remaining ones walk a Gray code (one XOR per further group).
template <typename C, typename Real>
TCB_HD void lround_thread(unsigned char* tile, const LRound& R, const ME<Real>* mats, uint32_t tid, int tb, int ngb) {
    uint32_t b = R.d;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (i < tb) b ^= (0u - ((tid >> i) & 1u)) & R.gcol[i];
    const uint32_t nit = 1u << (ngb - tb);
    const bool vec = sizeof(C) == 8 && ((R.ncodes >> 8) & 1u);
#pragma unroll 1
    for (uint32_t it = 0; it < nit; ++it) {
        if (it > 0) {
            int z = 0;
            while (!((it >> z) & 1u)) ++z;
            b ^= R.gcol[tb + z];
        }
        lround_group<C, Real>(tile, b, R, mats, vec);
    }
}

// ---- write-back through the final index map ----------------------------------------------------
// Same global addressing as stage_out (common.cuh); the shared-memory side reads logical element
// e = u * APU (and e + 1) at  d ^ XOR_t e_t col[t].
template <typename C>
TCB_HD void lstage_out_thread(const TileGeom& g, C* vec, uint64_t base, const unsigned char* tile, const uint64_t* rowoff,
                              const LOut& L, int tid, int nthr, int tb) {
    constexpr int APU = 16 / (int)sizeof(C);
    constexpr int S = APU == 2 ? 1 : 0;  // logical bit of unit-index bit 0
    const int ub = g.T - S;              // unit-index bits
    const uint32_t nunits = 1u << ub;
    const uint32_t rowmask = (1u << g.lrow) - 1u;
    if ((uint32_t)tid >= nunits) return;
    (void)nthr;
    uint32_t a = L.d;
    for (int i = 0; i < tb && i < ub; ++i) a ^= (0u - (((uint32_t)tid >> i) & 1u)) & L.col[i + S];
    const uint32_t nit = ub > tb ? (1u << (ub - tb)) : 1u;
    for (uint32_t it = 0; it < nit; ++it) {
        if (it > 0) {
            int z = 0;
            while (!((it >> z) & 1u)) ++z;
            // Gray code over the iteration bits: the visited unit is tid | gray(it) << tb
            a ^= L.col[tb + z + S];
        }
        const uint32_t gray = it ^ (it >> 1);
        const uint32_t uu = (uint32_t)tid | (gray << tb);
        const uint32_t e = uu * APU;
        const uint64_t gi = base + rowoff[e >> g.lrow] + (e & rowmask);
        Unit16 q;
        if (APU == 1 || L.vec) {
            q = *reinterpret_cast<const Unit16*>(tile + a);
        } else {
            C* qc = reinterpret_cast<C*>(&q);
            qc[0] = *reinterpret_cast<const C*>(tile + a);
            qc[1 % APU] = *reinterpret_cast<const C*>(tile + (a ^ L.col[0]));
        }
        *reinterpret_cast<Unit16*>(vec + gi) = q;
    }
}

// ---- production shape: 256 threads, compile-time loop structure ---------------------------------
// NIT = groups per thread per round (2 for the 2^13-amplitude complex64 tile, 1 for complex128)
template <typename C, typename Real, int NIT>
TCB_HD void lround_thread_fast(unsigned char* tile, const LRound& R, const ME<Real>* mats, uint32_t tid) {
    uint32_t b = R.d;
#pragma unroll
    for (int i = 0; i < 8; ++i) b ^= (0u - ((tid >> i) & 1u)) & R.gcol[i];
    const bool vec = sizeof(C) == 8 && ((R.ncodes >> 8) & 1u);
#pragma unroll 1
    for (int it = 0; it < NIT; ++it) {
        if (it > 0) b ^= R.gcol[8];
        lround_group<C, Real>(tile, b, R, mats, vec);
    }
}

// amplitude offset (relative to the tile base) of the thread's first element tid * APU
template <int APU>
TCB_HD uint64_t lstage_thread_goff(const LStage& S, uint32_t tid) {
    const uint32_t e = tid * APU;
    uint64_t o = 0;
#pragma unroll
    for (int b = 0; b < 9; ++b) o += ((e >> b) & 1u) ? S.bit_off[b] : 0ull;
    return o;
}

template <typename C>
TCB_HD void lstage_in_fast(const LStage& S, const C* vec_base, unsigned char* tile, uint32_t tid) {
    constexpr int APU = 16 / (int)sizeof(C);
    const C* g = vec_base + lstage_thread_goff<APU>(S, tid);
    const uint32_t s0 = swz_unit(tid) << 4;
#pragma unroll
    for (int i = 0; i < LP_FAST_ITERS; ++i) {
#if defined(__CUDA_ARCH__)
        cp_async16(tile + (s0 ^ S.sin[i]), g + S.goff[i]);
#else
        *reinterpret_cast<Unit16*>(tile + (s0 ^ S.sin[i])) = *reinterpret_cast<const Unit16*>(g + S.goff[i]);
#endif
    }
}

template <typename C>
TCB_HD void lstage_out_fast(const LStage& S, const LOut& L, C* vec_base, const unsigned char* tile, uint32_t tid) {
    constexpr int APU = 16 / (int)sizeof(C);
    constexpr int SH = APU == 2 ? 1 : 0;
    C* g = vec_base + lstage_thread_goff<APU>(S, tid);
    uint32_t a = L.d;
#pragma unroll
    for (int i = 0; i < 8; ++i) a ^= (0u - ((tid >> i) & 1u)) & L.col[i + SH];
    const bool vec = APU == 1 || L.vec;
    const uint32_t c0 = L.col[0];
#pragma unroll
    for (int i = 0; i < LP_FAST_ITERS; ++i) {
        const uint32_t ai = a ^ S.sout[i];
        Unit16 q;
        if (vec) {
            q = *reinterpret_cast<const Unit16*>(tile + ai);
        } else {
            C* qc = reinterpret_cast<C*>(&q);
            qc[0] = *reinterpret_cast<const C*>(tile + ai);
            qc[1 % APU] = *reinterpret_cast<const C*>(tile + (ai ^ c0));
        }
        *reinterpret_cast<Unit16*>(g + S.goff[i]) = q;
    }
}

// ---- host side (lpass.cu) ----------------------------------------------------------------------
struct LPassInfo {
    int rounds;          // shared-memory round trips
    int nlin;            // gates absorbed into the index map (no device work)
    int ndiag;           // diagonal gates (table multiplies)
    int ndense;          // dense gates
    int conflicts;       // rounds whose accesses are not bank-conflict-free
    int vec_rounds;      // rounds using 16-byte accesses
    double fma_per_amp;  // real FMAs per amplitude for the whole pass
    int mat_elems;       // parameter-bank elements used
};

}  // namespace tcb
