// CPU emulation of the tcb200 kernels -- TEST INFRASTRUCTURE ONLY (the build container has no
// GPU).  It runs the *same* __host__ __device__ bodies the kernels run (stage_in,
// apply_block_on_tile, stage_out, expect_tile_term, tile_base, row_offset, the host-side
// make_geom / make_group_map planning) thread by thread and phase by phase, so that index,
// swizzle and planning logic is checked against the oracle before any GPU time is spent.
// The product library never links this file.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"

using namespace tcb;

namespace {

template <typename Real, int K>
int emu_dense_t(void* state, int nbits, const int* bits, const double* mat) {
    using C = typename CT<Real>::type;
    constexpr int D = 1 << K;
    constexpr int APU = CT<Real>::APU;
    TileGeom g;
    GroupMap gm;
    const int tile_bits = dense_tile_bits(sizeof(Real) == 4 ? TCB200_C64 : TCB200_C128, K);
    int rc = make_geom(nbits, tile_bits, K, bits, &g);
    if (rc) return rc;
    rc = make_group_map(g, APU, K, bits, &gm);
    if (rc) return rc;
    const int tb = pick_threads(g.T, K, APU);
    const int nthr = 1 << tb;
    std::vector<C> m(D * D);
    for (int i = 0; i < D * D; ++i) {
        m[i].x = (Real)mat[2 * i];
        m[i].y = (Real)mat[2 * i + 1];
    }
    const size_t tile_elems = (size_t)1 << g.T;
    C* tile = static_cast<C*>(aligned_alloc(128, tile_elems * sizeof(C) < 128 ? 128 : tile_elems * sizeof(C)));
    uint64_t rowoff[256];
    for (int r = 0; r < (1 << g.h); ++r) rowoff[r] = row_offset(g, r);
    C* vec = static_cast<C*>(state);
    const uint64_t ntiles = 1ull << (nbits - g.T);
    for (uint64_t t = 0; t < ntiles; ++t) {
        const uint64_t base = tile_base(g, t);
        for (int tid = 0; tid < nthr; ++tid) stage_in<C, true>(g, vec, base, tile, rowoff, tid, nthr);
        for (int tid = 0; tid < nthr; ++tid)
            apply_block_dispatch<C,K>(tile, gm, tid, nthr, tb, [&](int i, int j) { return m[i * D + j]; });
        for (int tid = 0; tid < nthr; ++tid) stage_out<C, true>(g, vec, base, tile, rowoff, tid, nthr);
    }
    free(tile);
    return 0;
}

template <typename Real>
int emu_dense_k(void* state, int nbits, int k, const int* bits, const double* mat) {
    switch (k) {
        case 1: return emu_dense_t<Real, 1>(state, nbits, bits, mat);
        case 2: return emu_dense_t<Real, 2>(state, nbits, bits, mat);
        case 3: return emu_dense_t<Real, 3>(state, nbits, bits, mat);
        case 4: return emu_dense_t<Real, 4>(state, nbits, bits, mat);
        case 5: return emu_dense_t<Real, 5>(state, nbits, bits, mat);
    }
    return -2;
}

template <typename Real>
int emu_pass_t(void* state, int nbits, int nops, const int* ops_k, const int* ops_bits, const double* mats,
               int n_hi, const int* tile_hi) {
    using C = typename CT<Real>::type;
    constexpr int APU = CT<Real>::APU;
    TileGeom g;
    const int tile_bits = pass_tile_bits(sizeof(Real) == 4 ? TCB200_C64 : TCB200_C128);
    int rc = make_geom_hi(nbits, tile_bits, nbits <= tile_bits ? 0 : n_hi, tile_hi, &g);
    if (rc) return rc;
    std::vector<GroupMap> gms(nops);
    std::vector<std::vector<C>> ms(nops);
    const int* b = ops_bits;
    const double* mp = mats;
    int kmin = 99;
    for (int o = 0; o < nops; ++o) {
        const int k = ops_k[o];
        rc = make_group_map(g, APU, k, b, &gms[o]);
        if (rc) return rc;
        const int sz = 1 << (2 * k);
        ms[o].resize(sz);
        for (int i = 0; i < sz; ++i) {
            ms[o][i].x = (Real)mp[2 * i];
            ms[o][i].y = (Real)mp[2 * i + 1];
        }
        b += k;
        mp += 2 * sz;
        if (k < kmin) kmin = k;
    }
    const int tb = pick_threads(g.T, kmin, APU);
    const int nthr = 1 << tb;
    const size_t tile_elems = (size_t)1 << g.T;
    C* tile = static_cast<C*>(aligned_alloc(128, tile_elems * sizeof(C) < 128 ? 128 : tile_elems * sizeof(C)));
    uint64_t rowoff[256];
    for (int r = 0; r < (1 << g.h); ++r) rowoff[r] = row_offset(g, r);
    C* vec = static_cast<C*>(state);
    const uint64_t ntiles = 1ull << (nbits - g.T);
    for (uint64_t t = 0; t < ntiles; ++t) {
        const uint64_t base = tile_base(g, t);
        for (int tid = 0; tid < nthr; ++tid) stage_in<C, true>(g, vec, base, tile, rowoff, tid, nthr);
        for (int o = 0; o < nops; ++o) {
            const C* m = ms[o].data();
            for (int tid = 0; tid < nthr; ++tid) {
                switch (ops_k[o]) {
                    case 1: apply_block_dispatch<C,1>(tile, gms[o], tid, nthr, tb, [&](int i, int j) { return m[i * 2 + j]; }); break;
                    case 2: apply_block_dispatch<C,2>(tile, gms[o], tid, nthr, tb, [&](int i, int j) { return m[i * 4 + j]; }); break;
                    case 3: apply_block_dispatch<C,3>(tile, gms[o], tid, nthr, tb, [&](int i, int j) { return m[i * 8 + j]; }); break;
                    default: apply_block_dispatch<C,4>(tile, gms[o], tid, nthr, tb, [&](int i, int j) { return m[i * 16 + j]; }); break;
                }
            }
        }
        for (int tid = 0; tid < nthr; ++tid) stage_out<C, true>(g, vec, base, tile, rowoff, tid, nthr);
    }
    free(tile);
    return 0;
}

template <typename Real>
int emu_expect_t(const void* state, int nbits, int nterms, const uint64_t* flip, const uint64_t* sign,
                 const int* ny, int n_hi, const int* tile_hi, double* out) {
    using C = typename CT<Real>::type;
    TileGeom g;
    const int T = expect_tile_bits(sizeof(Real) == 4 ? TCB200_C64 : TCB200_C128);
    int rc = make_geom_hi(nbits, T, nbits <= T ? 0 : n_hi, tile_hi, &g);
    if (rc) return rc;
    const size_t tile_elems = (size_t)1 << g.T;
    C* tile = static_cast<C*>(aligned_alloc(128, tile_elems * sizeof(C) < 128 ? 128 : tile_elems * sizeof(C)));
    uint64_t rowoff[256];
    for (int r = 0; r < (1 << g.h); ++r) rowoff[r] = row_offset(g, r);
    const C* vec = static_cast<const C*>(state);
    const uint64_t ntiles = 1ull << (nbits - g.T);
    const int nthr = 32;
    for (int t = 0; t < nterms; ++t) {
        uint32_t fl = 0, sl = 0;
        uint64_t shi = sign[t];
        for (int b = 0; b < nbits; ++b) {
            const int lb = local_bit(g, b);
            if ((flip[t] >> b) & 1ull) {
                if (lb < 0) return fail(TCB200_ERR_ARG, "flip bit outside the tile");
                fl |= 1u << lb;
            }
            if (((sign[t] >> b) & 1ull) && lb >= 0) {
                sl |= 1u << lb;
                shi &= ~(1ull << b);
            }
        }
        double re = 0, im = 0;
        for (uint64_t tl = 0; tl < ntiles; ++tl) {
            const uint64_t base = tile_base(g, tl);
            for (int tid = 0; tid < nthr; ++tid) stage_in<C, false>(g, vec, base, tile, rowoff, tid, nthr);
            for (int tid = 0; tid < nthr; ++tid) {
                // the kernel's multi-term body, run here with this term in slot 0
                Real pv[TCB200_MAX_TERMS];
                const uint32_t fla[TCB200_MAX_TERMS] = {fl}, sla[TCB200_MAX_TERMS] = {sl};
                expect_tile_terms<C, Real, TCB200_MAX_TERMS>(tile, (uint32_t)tile_elems, 1, fla, sla, (uint32_t)(ny[t] & 1), tid, nthr, pv);
                const bool neg = parity64(base & shi);
                const double v = neg ? -(double)pv[0] : (double)pv[0];
                if (ny[t] & 1) im += v;
                else re += v;
            }
        }
        double ore = re, oim = im;
        switch (ny[t] & 3) {
            case 1: ore = im; oim = -re; break;
            case 2: ore = -re; oim = -im; break;
            case 3: ore = -im; oim = re; break;
            default: break;
        }
        out[2 * t] = ore;
        out[2 * t + 1] = oim;
    }
    free(tile);
    return 0;
}

}  // namespace

#define EMU_API extern "C" __attribute__((visibility("default")))

EMU_API int emu_apply_dense(void* state, int nbits, int dtype, int k, const int* bits, const double* mat) {
    return dtype == TCB200_C64 ? emu_dense_k<float>(state, nbits, k, bits, mat)
                               : emu_dense_k<double>(state, nbits, k, bits, mat);
}

EMU_API int emu_apply_pass(void* state, int nbits, int dtype, int nops, const int* ops_k, const int* ops_bits,
                           const double* mats, int n_hi, const int* tile_hi) {
    return dtype == TCB200_C64 ? emu_pass_t<float>(state, nbits, nops, ops_k, ops_bits, mats, n_hi, tile_hi)
                               : emu_pass_t<double>(state, nbits, nops, ops_k, ops_bits, mats, n_hi, tile_hi);
}

EMU_API int emu_expect(const void* state, int nbits, int dtype, int nterms, const uint64_t* flip,
                       const uint64_t* sign, const int* ny, int n_hi, const int* tile_hi, double* out) {
    return dtype == TCB200_C64 ? emu_expect_t<float>(state, nbits, nterms, flip, sign, ny, n_hi, tile_hi, out)
                               : emu_expect_t<double>(state, nbits, nterms, flip, sign, ny, n_hi, tile_hi, out);
}

EMU_API int emu_pass_tile_bits(int dtype) { return pass_tile_bits(dtype); }
EMU_API int emu_expect_tile_bits(int dtype) { return expect_tile_bits(dtype); }
EMU_API int emu_dense_tile_bits(int dtype, int k) { return dense_tile_bits(dtype, k); }
EMU_API const char* emu_last_error() { return tcb200_last_error(); }

// bank-conflict audit of a group map: worst number of distinct 16-byte-bank-group addresses that
// collide inside one shared-memory access phase (1 = conflict free)
EMU_API int emu_conflict_degree(int nbits, int dtype, int k, const int* bits) {
    TileGeom g;
    GroupMap gm;
    const int apu = dtype == TCB200_C64 ? 2 : 1;
    if (make_geom(nbits, dense_tile_bits(dtype, k), k, bits, &g)) return -1;
    if (make_group_map(g, apu, k, bits, &gm)) return -1;
    const int tb = pick_threads(g.T, k, apu);
    const int esz = dtype == TCB200_C64 ? 8 : 16;
    const bool vec = gm.vec0 || esz == 16;      // 16-byte accesses: 8 lanes per phase
    const int lanes_per_phase = vec ? 8 : 16;   // 8-byte accesses: 16 lanes per phase
    int worst = 1;
    const uint32_t ngroups = 1u << gm.ngb;
    const uint32_t nthr = 1u << tb;
    for (uint32_t w0 = 0; w0 < nthr && w0 < ngroups; w0 += lanes_per_phase) {
        for (uint32_t j = 0; j < (1u << k); ++j) {
            int cnt[32] = {0};
            for (int l = 0; l < lanes_per_phase; ++l) {
                const uint32_t gi = w0 + l;
                if (gi >= ngroups) break;
                const uint32_t e = group_base(gm, gi) ^ gm.tval[j];
                const uint32_t byte = e * esz;
                const int bank = vec ? (byte / 16) % 8 : (byte / 8) % 16;
                cnt[bank]++;
            }
            for (int b = 0; b < 32; ++b)
                if (cnt[b] > worst) worst = cnt[b];
        }
    }
    return worst;
}
