// Error plumbing, version, launch counter and the host-side tile / lane planning shared by
// the kernels.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace tcb {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char* what) {
    return fail(TCB200_ERR_CUDA - (int)e, "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e),
                what);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// -------------------------------------------------------------------------------------------
int make_geom_hi(int nbits, int tile_bits, int n_hi, const int* tile_hi, TileGeom* g, int max_hi) {
    if (nbits < 1 || nbits > 62) return fail(TCB200_ERR_ARG, "nbits=%d out of range", nbits);
    if (n_hi < 0 || n_hi > max_hi) return fail(TCB200_ERR_ARG, "n_hi=%d out of range (max %d)", n_hi, max_hi);
    memset(g, 0, sizeof(*g));
    g->n = nbits;
    if (nbits <= tile_bits) {  // whole vector is one tile
        g->T = nbits;
        g->h = 0;
        g->lrow = nbits;
        return 0;
    }
    g->T = tile_bits;
    g->h = n_hi;
    g->lrow = tile_bits - n_hi;
    if (g->lrow < 1) return fail(TCB200_ERR_ARG, "too many gathered bits for the tile");
    for (int j = 0; j < n_hi; ++j) {
        const int b = tile_hi[j];
        if (b < g->lrow || b >= nbits || (j > 0 && b <= tile_hi[j - 1]))
            return fail(TCB200_ERR_ARG, "gathered bit %d invalid (lrow=%d, nbits=%d)", b, g->lrow,
                        nbits);
        g->hb[j] = b;
    }
    return 0;
}

int make_geom(int nbits, int tile_bits, int k, const int* bits, TileGeom* g) {
    for (int i = 0; i < k; ++i) {
        if (bits[i] < 0 || bits[i] >= nbits)
            return fail(TCB200_ERR_ARG, "bit %d out of range for a %d-bit state", bits[i], nbits);
        if (i > 0 && bits[i] <= bits[i - 1])
            return fail(TCB200_ERR_ARG, "bits must be strictly ascending");
    }
    if (nbits <= tile_bits) return make_geom_hi(nbits, tile_bits, 0, nullptr, g);
    // h = number of targets at or above the contiguous part [0, tile_bits - h)
    int h = 0;
    for (;;) {
        int c = 0;
        for (int i = 0; i < k; ++i)
            if (bits[i] >= tile_bits - h) ++c;
        if (c == h) break;
        h = c;
    }
    int hi[8];
    int nh = 0;
    for (int i = 0; i < k; ++i)
        if (bits[i] >= tile_bits - h) hi[nh++] = bits[i];
    return make_geom_hi(nbits, tile_bits, nh, hi, g);
}

// bank class of a local amplitude bit under the tile swizzle (see common.cuh): bits that end up
// in the same class move the same bank-select bit.  SWZ_SW folds two 3-bit fields onto the
// chunk index, SWZ_HW128 (the TMA pattern) one.
static int bank_class(int apu, int b, int mode) {
    const int top = (mode == SWZ_HW128) ? 6 : 9;   // highest unit bit + 1 that still moves a bank bit
    if (apu == 2) {          // complex64: bit 0 = which half of the 16-byte unit
        if (b == 0) return 0;
        if (b <= top) return 1 + (b - 1) % 3;
        return -1;
    }
    if (b < top) return b % 3;  // complex128: one amplitude per unit
    return -1;
}

int make_group_map(const TileGeom& g, int apu, int k, const int* bits, GroupMap* gm, int swz_mode) {
    memset(gm, 0, sizeof(*gm));
    if (k < 1 || k > TCB200_MAX_K) return fail(TCB200_ERR_UNSUPPORTED, "k=%d unsupported", k);
    if (k > g.T) return fail(TCB200_ERR_ARG, "k=%d exceeds the state size", k);
    int tl[8];
    bool is_t[32] = {false};
    for (int i = 0; i < k; ++i) {
        tl[i] = local_bit(g, bits[i]);
        if (tl[i] < 0) return fail(TCB200_ERR_ARG, "bit %d is not inside the tile", bits[i]);
        if (i > 0 && bits[i] <= bits[i - 1])
            return fail(TCB200_ERR_ARG, "bits must be strictly ascending");
        is_t[tl[i]] = true;
    }
    gm->ngb = g.T - k;
    if (gm->ngb > 16) return fail(TCB200_ERR_UNSUPPORTED, "tile too large for the group map");
    gm->vec0 = (apu == 2 && tl[0] == 0) ? 1 : 0;
    // lane order of the non-target bits: one representative per bank class first
    int order[32];
    int no = 0;
    bool used[32] = {false};
    const int ncls = (apu == 2) ? 4 : 3;
    for (int c = (gm->vec0 ? 1 : 0); c < ncls; ++c) {
        for (int b = 0; b < g.T; ++b) {
            if (!is_t[b] && !used[b] && bank_class(apu, b, swz_mode) == c) {
                order[no++] = b;
                used[b] = true;
                break;
            }
        }
    }
    for (int b = 0; b < g.T; ++b)
        if (!is_t[b] && !used[b]) order[no++] = b;
    for (int i = 0; i < gm->ngb; ++i)
        gm->ntval[i] = (apu == 2) ? swz_amp<2>(1u << order[i], swz_mode) : swz_amp<1>(1u << order[i], swz_mode);
    for (uint32_t j = 0; j < (1u << k); ++j) {
        const uint32_t e = deposit(j, tl, k);
        gm->tval[j] = (apu == 2) ? swz_amp<2>(e, swz_mode) : swz_amp<1>(e, swz_mode);
    }
    return 0;
}

int pick_threads(int T, int k, int apu) {
    // one 16-byte unit per lane per staging iteration; enough groups to keep every lane busy
    int groups_log = T - k;
    int tb = groups_log < 8 ? groups_log : 8;
    int units_log = T - (apu == 2 ? 1 : 0);
    if (tb > units_log) tb = units_log;
    if (tb < 5) tb = 5;
    return tb;
}

static int env_tile_log2(const char* name, int dflt) {
    const char* e = getenv(name);
    if (e) {
        const int v = atoi(e);
        if (v >= 7 && v <= 17) return v;
    }
    return dflt;
}

int dense_tile_bits(int dtype, int k) {
    // 32 KiB tiles: 4096 complex64 / 2048 complex128 (measured alternatives in DESIGN.md);
    // TCB200_TILE_BYTES_LOG2 overrides (used by the tests to exercise many tiles at small n).
    const int bytes_log2 = env_tile_log2("TCB200_TILE_BYTES_LOG2", 15);
    int t = bytes_log2 - (dtype == TCB200_C64 ? 3 : 4);
    if (k == 5 && t - k < 7 && bytes_log2 >= 15) t = k + 7;  // keep >= 128 groups per tile
    if (t < k + 1) t = k + 1;
    return t;
}

int pass_tile_bits(int dtype) {
    // 64 KiB tiles: three CTAs of cpass_kernel per SM; measured alternatives in DESIGN.md
    const int bytes_log2 = env_tile_log2("TCB200_PASS_TILE_BYTES_LOG2", PASS_TILE_BYTES_LOG2);
    return bytes_log2 - (dtype == TCB200_C64 ? 3 : 4);
}

int expect_tile_bits(int dtype) {
    const int bytes_log2 = env_tile_log2("TCB200_EXPECT_TILE_BYTES_LOG2", 15);
    return bytes_log2 - (dtype == TCB200_C64 ? 3 : 4);
}

}  // namespace tcb

extern "C" {

const char* tcb200_version(void) { return "tcb200 0.1.0 sm_100a"; }

const char* tcb200_last_error(void) { return tcb::g_err; }

int64_t tcb200_launch_count(void) { return tcb::g_launches.load(); }

}  // extern "C"
