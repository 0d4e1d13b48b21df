// Single-flip Pauli strings -- X_j or Y_j with any Z dressing: the off-diagonal half of every
// transverse-field / Heisenberg-type Hamiltonian -- evaluated 12 per read of the state.
//
// Replaces expectation_before + contraction (tensorcircuit/basecircuit.py:267-319,
// circuit.py:914-990) for these strings, like tcb200_expect_pauli, but with the register-tile
// structure of the gate pass instead of one shared-memory partner load per amplitude and string:
// a CTA stages a 64 KiB tile (the gate pass's production shape and staging code), and in each of
// up to three rounds a thread loads 16 amplitudes -- 4 tile bits -- once and forms, for each of
// those 4 bits j, the pair sums
//     sum_{e: e_j = 0} s(e) * Re / Im( conj(psi_e) psi_{e | 1<<j} )         (quantum.py:1461-1482)
// entirely in registers (2 FMA per pair).  12 flip bits, one string each, per launch.  Signs of the Z
// dressing split into a per-register mask, a per-thread parity and a per-tile parity.  Per-thread
// partial sums are kept in the state's real type for 16 tiles at a time and then folded into
// float64 (warp shuffle -> shared memory), so the result does not depend on the size of the state.
#include <string.h>

#include <algorithm>
#include <vector>

#include "lpass.cuh"

namespace tcb {

constexpr int XE_ROUNDS = 3;
constexpr int XE_SLOTS = 1;                                // strings per flip bit and launch
constexpr int XE_TERMS = XE_ROUNDS * LP_RB * XE_SLOTS;     // 24
constexpr int XE_FLUSH = 16;                               // tiles between float -> double folds

struct XESlot {
    int32_t term;       // output index, -1: empty
    uint32_t imag;      // 1: the Y-type component  a0.x a1.y - a0.y a1.x
    uint32_t negmask;   // bit k0 (register index with the flip position cleared): negate that pair
    uint32_t gmask;     // parity of (group index & gmask) negates the group's sum
    uint64_t himask;    // parity of (tile base & himask) negates the tile's sum
};

struct XEParams {
    const void* state;
    double* partials;  // [batch][gridDim.x][XE_TERMS]
    TileGeom g;
    LStage stage;
    int nrounds;
    uint32_t gcol[XE_ROUNDS][LP_MAX_GB];
    uint32_t rcol[XE_ROUNDS][LP_RB];
    XESlot slot[XE_ROUNDS][LP_RB][XE_SLOTS];
};

template <typename C, typename Real, int P, bool IMAG, bool SIGNED>
__device__ __forceinline__ Real xe_pairs_t(const C* v, const uint32_t negmask) {
    Real tot = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int lo = r & ((1 << P) - 1);
        const int i0 = ((r >> P) << (P + 1)) | lo, i1 = i0 | (1 << P);
        Real d;
        if (IMAG) {
            d = v[i0].x * v[i1].y;
            d = fma(-v[i0].y, v[i1].x, d);
        } else {
            d = v[i0].x * v[i1].x;
            d = fma(v[i0].y, v[i1].y, d);
        }
        if (SIGNED) tot += ((negmask >> i0) & 1u) ? -d : d;
        else tot += d;
    }
    return tot;
}

// the four code variants are selected by warp-uniform flags of the slot
template <typename C, typename Real, int P>
__device__ __forceinline__ Real xe_pairs(const C* v, const XESlot& s) {
    if (s.negmask == 0u) return s.imag ? xe_pairs_t<C, Real, P, true, false>(v, 0u) : xe_pairs_t<C, Real, P, false, false>(v, 0u);
    return s.imag ? xe_pairs_t<C, Real, P, true, true>(v, s.negmask) : xe_pairs_t<C, Real, P, false, true>(v, s.negmask);
}

template <typename Real>
__global__ void __launch_bounds__(256, sizeof(Real) == 4 ? 3 : 2) xexpect_kernel(const __grid_constant__ XEParams p) {
    using C = typename CT<Real>::type;
    constexpr int NIT = sizeof(C) == 8 ? 2 : 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double dacc[8][XE_TERMS];
    const uint32_t tid = threadIdx.x;
    const int w = tid >> 5, l = tid & 31;
    for (int i = tid; i < 8 * XE_TERMS; i += 256) (&dacc[0][0])[i] = 0.0;
    const C* vec0 = static_cast<const C*>(p.state) + ((uint64_t)blockIdx.y << p.g.n);
    const uint64_t ntiles = 1ull << (p.g.n - p.g.T);

    Real acc[XE_ROUNDS][LP_RB][XE_SLOTS];
#pragma unroll
    for (int r = 0; r < XE_ROUNDS; ++r)
#pragma unroll
        for (int j = 0; j < LP_RB; ++j)
#pragma unroll
            for (int q = 0; q < XE_SLOTS; ++q) acc[r][j][q] = 0;

    auto fold = [&]() {  // float partials -> float64, fixed order
#pragma unroll
        for (int r = 0; r < XE_ROUNDS; ++r)
#pragma unroll
            for (int j = 0; j < LP_RB; ++j)
#pragma unroll
                for (int q = 0; q < XE_SLOTS; ++q) {
                    if (p.slot[r][j][q].term < 0) continue;
                    double x = (double)acc[r][j][q];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
                    if (l == 0) dacc[w][(r * LP_RB + j) * XE_SLOTS + q] += x;
                    acc[r][j][q] = 0;
                }
    };

    int since = 0;
    for (uint64_t tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
        const uint64_t base = tile_base(p.g, tl);
        __syncthreads();  // the previous tile is no longer read
        lstage_in_fast<C>(p.stage, vec0 + base, smem_raw, tid);
        cp_async_wait_all();
        __syncthreads();
#pragma unroll
        for (int r = 0; r < XE_ROUNDS; ++r) {
            if (r >= p.nrounds) break;
            uint32_t b = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) b ^= (0u - ((tid >> i) & 1u)) & p.gcol[r][i];
            const uint32_t c0 = p.rcol[r][0], c1 = p.rcol[r][1], c2 = p.rcol[r][2], c3 = p.rcol[r][3];
#pragma unroll 1
            for (int it = 0; it < NIT; ++it) {
                if (it > 0) b ^= p.gcol[r][8];
                const uint32_t gidx = tid | ((uint32_t)it << 8);
                C v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const uint32_t off = ((j & 1) ? c0 : 0u) ^ ((j & 2) ? c1 : 0u) ^ ((j & 4) ? c2 : 0u) ^ ((j & 8) ? c3 : 0u);
                    v[j] = *reinterpret_cast<const C*>(smem_raw + (b ^ off));
                }
#pragma unroll
                for (int j = 0; j < LP_RB; ++j)
#pragma unroll
                    for (int q = 0; q < XE_SLOTS; ++q) {
                        const XESlot& s = p.slot[r][j][q];
                        if (s.term < 0) continue;
                        Real t;
                        if (j == 0) t = xe_pairs<C, Real, 0>(v, s);
                        else if (j == 1) t = xe_pairs<C, Real, 1>(v, s);
                        else if (j == 2) t = xe_pairs<C, Real, 2>(v, s);
                        else t = xe_pairs<C, Real, 3>(v, s);
                        const uint32_t par = (__popc(gidx & s.gmask) + __popcll(base & s.himask)) & 1u;
                        acc[r][j][q] += par ? -t : t;
                    }
            }
        }
        if (++since == XE_FLUSH) {
            fold();
            since = 0;
        }
    }
    fold();
    __syncthreads();
    if (tid < XE_TERMS) {
        double s = 0.0;
        for (int ww = 0; ww < 8; ++ww) s += dacc[ww][tid];
        p.partials[((uint64_t)blockIdx.y * gridDim.x + blockIdx.x) * XE_TERMS + tid] = 2.0 * s;  // both orders of every pair
    }
}

struct XEFinalParams {
    const double* partials;
    double* out;  // [batch][nterms][2]
    int nctas;
    int nterms;
    int slot_of[XE_TERMS];  // term -> accumulator index
    int ny[XE_TERMS];
};

// grid (batch, nterms): fixed-order sum over the CTAs, then the (-i)^ny phase (expect.cu does the same)
__global__ void __launch_bounds__(256) xexpect_final_kernel(const __grid_constant__ XEFinalParams p) {
    __shared__ double red[8];
    const int t = blockIdx.y, tid = threadIdx.x;
    const double* src = p.partials + (uint64_t)blockIdx.x * p.nctas * XE_TERMS + p.slot_of[t];
    double s = 0.0;
    for (int c = tid; c < p.nctas; c += 256) s += src[(uint64_t)c * XE_TERMS];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((tid & 31) == 0) red[tid >> 5] = s;
    __syncthreads();
    if (tid != 0) return;
    s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w];
    // X-type: the sum is the real part; Y-type (ny = 1): it is the imaginary part, times (-i)
    double re = (p.ny[t] & 1) ? 0.0 : s, im = (p.ny[t] & 1) ? s : 0.0;
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

static unsigned xe_grid_x(int nbits, int T, int64_t batch) {
    const uint64_t ntiles = nbits > T ? (1ull << (nbits - T)) : 1ull;
    uint64_t cap = (148ull * 3) / (uint64_t)(batch < 1 ? 1 : batch);
    if (cap < 4) cap = 4;
    return (unsigned)(ntiles < cap ? ntiles : cap);
}

}  // namespace tcb

using namespace tcb;

extern "C" {

int tcb200_expect_single_flip_max_terms(void) { return XE_TERMS; }

size_t tcb200_expect_single_flip_workspace_bytes(int nbits, int64_t batch) {
    const unsigned g0 = xe_grid_x(nbits, pass_tile_bits(TCB200_C64), batch), g1 = xe_grid_x(nbits, pass_tile_bits(TCB200_C128), batch);
    return (size_t)(batch < 1 ? 1 : batch) * (g0 > g1 ? g0 : g1) * XE_TERMS * sizeof(double) + 256;
}

int tcb200_expect_single_flip(const void* state, int nbits, int dtype, int nterms, const int* flip_bit, const uint64_t* sign,
                              const int* ny, int n_hi, const int* tile_hi, double* out_dev, int64_t batch, void* workspace,
                              size_t ws_bytes, void* stream) {
    if (!state || !flip_bit || !sign || !ny || !out_dev || !workspace) return fail(TCB200_ERR_ARG, "NULL argument");
    if (dtype != TCB200_C64 && dtype != TCB200_C128) return fail(TCB200_ERR_ARG, "bad dtype %d", dtype);
    if (nterms < 1 || nterms > XE_TERMS) return fail(TCB200_ERR_ARG, "nterms=%d out of range (max %d)", nterms, XE_TERMS);
    if (batch < 1 || batch > 65535) return fail(TCB200_ERR_ARG, "batch=%lld out of range", (long long)batch);
    const int T = pass_tile_bits(dtype);
    const int apu = dtype == TCB200_C64 ? 2 : 1;
    if (nbits <= T || T - (apu == 2 ? 1 : 0) != 12)
        return fail(TCB200_ERR_UNSUPPORTED, "single-flip expectation needs a state larger than one 64 KiB tile");
    if (ws_bytes < tcb200_expect_single_flip_workspace_bytes(nbits, batch)) return fail(TCB200_ERR_WORKSPACE, "workspace too small");
    static thread_local XEParams* pp = nullptr;
    if (!pp) pp = new XEParams();
    XEParams& p = *pp;
    memset(&p, 0, sizeof(p));
    int rc = make_geom_hi(nbits, T, n_hi, tile_hi, &p.g, 9);
    if (rc) return rc;
    p.state = state;
    p.partials = static_cast<double*>(workspace);
    // index map of the staged tile: the staging swizzle
    uint32_t col[LP_MAX_T];
    const uint32_t amp = apu == 2 ? 8u : 16u;
    for (int t = 0; t < p.g.T; ++t) col[t] = (apu == 2 ? swz_amp<2>(1u << t) : swz_amp<1>(1u << t)) * amp;
    {  // staging constants (the gate pass's production shape)
        auto goff_of = [&](uint32_t e) { return row_offset(p.g, e >> p.g.lrow) + (uint64_t)(e & ((1u << p.g.lrow) - 1u)); };
        for (int b = 0; b < 9; ++b) p.stage.bit_off[b] = goff_of(1u << b);
        for (int i = 0; i < LP_FAST_ITERS; ++i) {
            const uint32_t u = (uint32_t)i << 8;
            p.stage.goff[i] = goff_of(u * apu);
            p.stage.sin[i] = swz_unit(u) << 4;
            p.stage.sout[i] = 0;
        }
    }
    // distinct flip bits (tile-local), ascending; <= XE_SLOTS strings each
    std::vector<int> fbits;
    for (int t = 0; t < nterms; ++t) {
        if (flip_bit[t] < 0 || flip_bit[t] >= nbits) return fail(TCB200_ERR_ARG, "flip bit %d out of range", flip_bit[t]);
        const int lb = local_bit(p.g, flip_bit[t]);
        if (lb < 0) return fail(TCB200_ERR_ARG, "flip bit %d of term %d is not inside the tile", flip_bit[t], t);
        if (std::find(fbits.begin(), fbits.end(), lb) == fbits.end()) fbits.push_back(lb);
    }
    std::sort(fbits.begin(), fbits.end());
    if ((int)fbits.size() > XE_ROUNDS * LP_RB) return fail(TCB200_ERR_ARG, "more than %d distinct flip bits in one launch", XE_ROUNDS * LP_RB);
    p.nrounds = ((int)fbits.size() + LP_RB - 1) / LP_RB;
    XEFinalParams f;
    memset(&f, 0, sizeof(f));
    for (int r = 0; r < XE_ROUNDS; ++r)
        for (int j = 0; j < LP_RB; ++j)
            for (int q = 0; q < XE_SLOTS; ++q) p.slot[r][j][q].term = -1;
    for (int r = 0; r < p.nrounds; ++r) {
        // register bits of the round: its flip bits, filled up with other tile bits
        std::vector<int> R;
        for (int j = 0; j < LP_RB && r * LP_RB + j < (int)fbits.size(); ++j) R.push_back(fbits[r * LP_RB + j]);
        for (int t = p.g.T - 1; t >= 0 && (int)R.size() < LP_RB; --t)
            if (std::find(R.begin(), R.end(), t) == R.end()) R.push_back(t);
        std::sort(R.begin(), R.end());
        int gl[LP_MAX_T], ng = 0;
        for (int t = 0; t < p.g.T; ++t)
            if (std::find(R.begin(), R.end(), t) == R.end()) gl[ng++] = t;
        // lane order: bank-select bits first (8-byte accesses: offset bits 3..6; 16-byte: 4..6)
        const uint32_t bmask = apu == 2 ? 0x78u : 0x70u;
        const int need = apu == 2 ? 4 : 3;
        int order[LP_MAX_T], no = 0;
        bool used[LP_MAX_T] = {false};
        uint32_t basis[8];
        int nbz = 0;
        for (int i = 0; i < ng && no < need; ++i) {
            uint32_t x = col[gl[i]] & bmask;
            for (int k = 0; k < nbz; ++k)
                if ((x ^ basis[k]) < x) x ^= basis[k];
            if (x) {
                basis[nbz++] = x;
                std::sort(basis, basis + nbz, [](uint32_t a, uint32_t b) { return a > b; });
                order[no++] = gl[i];
                used[i] = true;
            }
        }
        for (int i = 0; i < ng; ++i)
            if (!used[i]) order[no++] = gl[i];
        for (int i = 0; i < ng; ++i) p.gcol[r][i] = col[order[i]];
        for (int j = 0; j < LP_RB; ++j) p.rcol[r][j] = col[R[j]];
        // strings of this round
        for (int t = 0; t < nterms; ++t) {
            const int lb = local_bit(p.g, flip_bit[t]);
            const int j = (int)(std::find(R.begin(), R.end(), lb) - R.begin());
            if (j >= LP_RB) continue;
            if ((int)(std::find(fbits.begin(), fbits.end(), lb) - fbits.begin()) / LP_RB != r) continue;
            int q = 0;
            while (q < XE_SLOTS && p.slot[r][j][q].term >= 0) ++q;
            if (q == XE_SLOTS) return fail(TCB200_ERR_ARG, "more than %d strings flip bit %d", XE_SLOTS, flip_bit[t]);
            XESlot& s = p.slot[r][j][q];
            s.term = t;
            s.imag = (ny[t] & 1) ? 1u : 0u;
            const bool y_here = (sign[t] >> flip_bit[t]) & 1ull;
            if (((ny[t] & 1) != 0) != y_here || ny[t] > 1) return fail(TCB200_ERR_ARG, "term %d is not a single-flip string", t);
            uint64_t sg = sign[t] & ~(1ull << flip_bit[t]);
            // split the Z dressing: register bits -> negmask, group bits -> gmask, outside -> himask
            uint32_t regmask = 0;
            s.gmask = 0;
            s.himask = 0;
            for (int b = 0; b < nbits; ++b) {
                if (!((sg >> b) & 1ull)) continue;
                const int tb = local_bit(p.g, b);
                if (tb < 0) {
                    s.himask |= 1ull << b;
                    continue;
                }
                const int rp = (int)(std::find(R.begin(), R.end(), tb) - R.begin());
                if (rp < LP_RB) regmask |= 1u << rp;
                else {
                    const int gp = (int)(std::find(order, order + ng, tb) - order);
                    s.gmask |= 1u << gp;
                }
            }
            s.negmask = 0;
            for (uint32_t k = 0; k < 16; ++k)
                if (__builtin_popcount(k & regmask) & 1) s.negmask |= 1u << k;
            f.slot_of[t] = (r * LP_RB + j) * XE_SLOTS + q;
            f.ny[t] = ny[t];
        }
    }
    const unsigned gx = xe_grid_x(nbits, T, batch);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    static bool attr = false;
    if (!attr) {
        TCB_CUDA(cudaFuncSetAttribute(xexpect_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
        TCB_CUDA(cudaFuncSetAttribute(xexpect_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
        attr = true;
    }
    dim3 grid(gx, (unsigned)batch);
    if (dtype == TCB200_C64) xexpect_kernel<float><<<grid, 256, 64 * 1024, st>>>(p);
    else xexpect_kernel<double><<<grid, 256, 64 * 1024, st>>>(p);
    TCB_LAUNCH_CHECK("xexpect_kernel");
    f.partials = p.partials;
    f.out = out_dev;
    f.nctas = (int)gx;
    f.nterms = nterms;
    dim3 fg((unsigned)batch, (unsigned)nterms);
    xexpect_final_kernel<<<fg, 256, 0, st>>>(f);
    TCB_LAUNCH_CHECK("xexpect_final_kernel");
    return 0;
}

}  // extern "C"
