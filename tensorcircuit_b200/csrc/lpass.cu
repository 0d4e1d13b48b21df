// Structure-aware staged pass: host-side gate classification, round scheduling and index-map
// tracking, the kernel, and the C ABI entry points (tcb200_apply_gate_pass / tcb200_gate_pass_info).
//
// Replaces, for a run of gates whose bits fit one tile, the per-gate tn.contract_between ->
// tensordot of tensorcircuit/cons.py:605-623 and, for the structured gates of
// tensorcircuit/gates.py:46-127, 826-865 (cnot / swap / x / cz / rzz / rz ...), the dense
// multiplication by a table multiply (diagonal) or by nothing at all (affine permutations).
#include <cuda.h>  // CUtensorMap (kernel parameter of the TMA-staged variant)
#include <stdlib.h>
#include <string.h>

#include <stdio.h>

#include <algorithm>
#include <chrono>
#include <complex>
#include <mutex>
#include <vector>

#include "lpass.cuh"

namespace tcb {

typedef std::complex<double> cd;

namespace {

// ------------------------------------------------------------------------------------------------
// classified gates on tile-local logical bits
// ------------------------------------------------------------------------------------------------
enum { GK_DENSE = 0, GK_DIAG = 1, GK_LIN = 2 };

struct GOp {
    int kind;
    int k;
    int bm;             // matrices held: 1 (shared by the batch) or the batch size (vmap)
    int lb[4];          // logical tile bit of matrix index bit j
    std::vector<cd> m;  // diag: D entries per held matrix (dense ops use `src`)
    const double* src = nullptr;  // dense: the caller's bm row-major D x D complex128 matrices (valid during the call)
    int linv[4];        // lin: L^{-1} e_i as k-bit masks (the inverse map is y -> L^{-1} y ^ cinv)
    int cinv;
    int half = 0;       // dense 1-bit gate whose elements are each real or imaginary: 1 = real diagonal +
                        // imaginary off-diagonal (rx-like), 2 = real matrix (h, ry): LOP_RA / LOP_RB
};

// split one input gate into classified ops (appended to `out`); < 0 on error.  `mat` holds bm
// matrices (bm > 1: one per vmap batch element); the class is that of the union of their
// non-zero patterns, so that one schedule serves the whole batch.
int classify(int k, const int* lb, const double* mat, int bm, std::vector<GOp>& out) {
    const int D = 1 << k;
    auto M = [&](int b, int i, int j) { return cd(mat[2 * (((size_t)b * D + i) * D + j)], mat[2 * (((size_t)b * D + i) * D + j) + 1]); };
    // union of the non-zero patterns of the bm matrices, in one sweep over the data (a vmap batch of 1024
    // diagonal gates would otherwise be scanned once per zero entry and check)
    bool nzp[16][16];
    for (int i = 0; i < D; ++i)
        for (int j = 0; j < D; ++j) nzp[i][j] = false;
    for (int b = 0; b < bm; ++b) {
        const double* mb = mat + 2 * (size_t)b * D * D;
        for (int i = 0; i < D; ++i)
            for (int j = 0; j < D; ++j)
                if (mb[2 * (i * D + j)] != 0.0 || mb[2 * (i * D + j) + 1] != 0.0) nzp[i][j] = true;
    }
    auto any_nz = [&](int i, int j) { return nzp[i][j]; };
    // monomial?  (exact zeros: gate matrices are built analytically on the host)
    std::vector<int> perm(D, -1);
    bool mono = true;
    std::vector<char> rowused(D, 0);
    for (int j = 0; j < D && mono; ++j) {
        int cnt = 0, at = -1;
        for (int i = 0; i < D; ++i)
            if (any_nz(i, j)) {
                ++cnt;
                at = i;
            }
        if (cnt != 1 || rowused[at]) mono = false;
        else {
            perm[j] = at;
            rowused[at] = 1;
        }
    }
    if (mono) {
        bool ident = true, allone = true;
        for (int j = 0; j < D; ++j) {
            if (perm[j] != j) ident = false;
            for (int b = 0; b < bm; ++b)
                if (M(b, perm[j], j) != cd(1, 0)) allone = false;
        }
        bool affine = false;
        int lcol[4] = {0, 0, 0, 0};
        const int c = perm[0];
        if (!ident && k <= 3) {
            affine = true;
            for (int i = 0; i < k; ++i) lcol[i] = perm[1 << i] ^ c;
            for (int j = 0; j < D && affine; ++j) {
                int img = c;
                for (int i = 0; i < k; ++i)
                    if ((j >> i) & 1) img ^= lcol[i];
                if (img != perm[j]) affine = false;
            }
        }
        if (ident || affine) {
            if (!allone) {
                if (k > 4) return fail(TCB200_ERR_UNSUPPORTED, "diagonal gate of %d bits inside a gate pass (max 4)", k);
                GOp g;
                g.kind = GK_DIAG;
                g.k = k;
                g.bm = bm;
                for (int i = 0; i < k; ++i) g.lb[i] = lb[i];
                g.m.resize((size_t)bm * D);
                for (int b = 0; b < bm; ++b)
                    for (int j = 0; j < D; ++j) g.m[(size_t)b * D + j] = M(b, perm[j], j);
                out.push_back(g);
            }
            if (!ident) {
                GOp g;
                g.kind = GK_LIN;
                g.k = k;
                g.bm = 1;
                for (int i = 0; i < k; ++i) g.lb[i] = lb[i];
                std::vector<int> inv(D);
                for (int j = 0; j < D; ++j) inv[perm[j]] = j;
                g.cinv = inv[0];
                for (int i = 0; i < k; ++i) g.linv[i] = inv[1 << i] ^ g.cinv;
                out.push_back(g);
            }
            return 0;
        }
    }
    if (k > 3) return fail(TCB200_ERR_UNSUPPORTED, "dense gate of %d bits inside a gate pass (max 3)", k);
    GOp g;
    g.kind = GK_DENSE;
    g.k = k;
    g.bm = bm;
    for (int i = 0; i < k; ++i) g.lb[i] = lb[i];
    g.src = mat;
    if (k == 1) {  // exact zeros of the real / imaginary parts (gate matrices are built analytically on the host)
        bool real_diag = true, imag_off = true, real_off = true;
        for (int b = 0; b < bm; ++b) {
            real_diag = real_diag && M(b, 0, 0).imag() == 0.0 && M(b, 1, 1).imag() == 0.0;
            imag_off = imag_off && M(b, 0, 1).real() == 0.0 && M(b, 1, 0).real() == 0.0;
            real_off = real_off && M(b, 0, 1).imag() == 0.0 && M(b, 1, 0).imag() == 0.0;
        }
        if (real_diag && real_off) g.half = 2;
        else if (real_diag && imag_off) g.half = 1;
    }
    out.push_back(g);
    return 0;
}

// rank of the vectors v[0..n) restricted to `mask` (GF(2))
int gf2_rank(const uint32_t* v, int n, uint32_t mask) {
    uint32_t basis[32];
    int r = 0;
    for (int i = 0; i < n; ++i) {
        uint32_t x = v[i] & mask;
        for (int j = 0; j < r; ++j)
            if ((x ^ basis[j]) < x) x ^= basis[j];
        if (x) {
            basis[r++] = x;
            // keep the basis reduced enough for the (x ^ b) < x test: sort descending
            std::sort(basis, basis + r, [](uint32_t a, uint32_t b) { return a > b; });
        }
    }
    return r;
}

template <typename Real>
struct LPassParams {
    typename CT<Real>::type* state;
    TileGeom g;
    int tb;      // log2(threads working on groups)
    int stb;     // log2(blockDim.x)
    int ngb;     // group-index bits: T - LP_RB
    int nrounds;
    int fast;    // production shape: 256 threads, 16 staging units per thread (LStage valid)
    const ME<Real>* bmats;  // vmap: DEVICE [batch][bstride] matrix elements (nullptr: shared matrices in m[])
    int bstride;
    TmaPlan tp;             // TMA-staged kernel: boxes of one tile
    int vtl;                // pipelined kernel: log2(tiles per state vector)
    uint64_t ntiles;        // pipelined kernel: tiles over all vectors of the batch
    LOut out;
    LStage stage;
    LRound r[LP_MAX_ROUNDS];
    ME<Real> m[LP_MAT_ELEMS];
};

template <typename Real>
void put_elem(ME<Real>& e, cd z);
template <>
void put_elem<float>(ME<float>& e, cd z) {
    e.re = e.pad = (float)z.real();  // (re, re, im, im): both FFMA2 operand pairs as loaded
    e.im0 = e.im1 = (float)z.imag();
}
template <>
void put_elem<double>(ME<double>& e, cd z) {
    e.re = z.real();
    e.im0 = z.imag();
}

// ------------------------------------------------------------------------------------------------
// round scheduler
// ------------------------------------------------------------------------------------------------
template <typename Real>
struct Scheduler {
    static constexpr int AMP = (int)sizeof(typename CT<Real>::type);
    static constexpr bool C64 = AMP == 8;
    int T = 0;
    uint32_t col[LP_MAX_T];
    uint32_t d = 0;
    std::vector<GOp> ops;
    std::vector<std::vector<int>> preds;
    std::vector<char> done;
    LPassParams<Real>* q = nullptr;
    LPassInfo info;
    int nmat = 0;
    int batch = 1;                 // > 1: per-element matrices go to `blob` ([batch][bstride])
    ME<Real>* blob = nullptr;
    size_t bstride = LP_MAT_ELEMS;

    // element `idx` of the parameter bank, for batch element b
    void put(int idx, int b, cd z) {
        if (blob) put_elem<Real>(blob[(size_t)b * bstride + idx], z);
        else put_elem<Real>(q->m[idx], z);
    }

    int swz_mode = SWZ_SW;  // SWZ_HW128: the layout a TMA tensor-map load with SWIZZLE_128B leaves behind

    void init_layout() {
        for (int t = 0; t < T; ++t) {
            const uint32_t e = 1u << t;
            col[t] = (C64 ? swz_amp<2>(e, swz_mode) : swz_amp<1>(e, swz_mode)) * (uint32_t)AMP;
        }
        d = 0;
    }

    // Dependencies by shared tile bits in program order, except that diagonal ops commute with each other: a
    // diagonal op is ordered only against the non-diagonal ops around it (an rzz ladder is not a chain), a
    // non-diagonal op against every diagonal op since the previous non-diagonal one on that bit.
    void build_deps() {
        const int n = (int)ops.size();
        preds.assign(n, {});
        int last[LP_MAX_T];                      // last non-diagonal op per bit
        std::vector<int> dsince[LP_MAX_T];       // diagonal ops after it
        for (int t = 0; t < LP_MAX_T; ++t) last[t] = -1;
        auto add = [&](int i, int p) {
            if (p >= 0 && std::find(preds[i].begin(), preds[i].end(), p) == preds[i].end()) preds[i].push_back(p);
        };
        for (int i = 0; i < n; ++i) {
            const bool dg = ops[i].kind == GK_DIAG;
            for (int j = 0; j < ops[i].k; ++j) {
                const int b = ops[i].lb[j];
                if (dg) {
                    add(i, last[b]);
                    dsince[b].push_back(i);
                } else {
                    if (!dsince[b].empty()) {
                        for (int p : dsince[b]) add(i, p);
                        dsince[b].clear();
                    } else {
                        add(i, last[b]);
                    }
                    last[b] = i;
                }
            }
        }
        done.assign(n, 0);
    }

    bool ready(int i) const {
        if (done[i]) return false;
        for (int p : preds[i])
            if (!done[p]) return false;
        return true;
    }

    void apply_lin(const GOp& g) {
        uint32_t nc[4];
        for (int i = 0; i < g.k; ++i) {
            uint32_t c = 0;
            for (int s = 0; s < g.k; ++s)
                if ((g.linv[i] >> s) & 1) c ^= col[g.lb[s]];
            nc[i] = c;
        }
        for (int s = 0; s < g.k; ++s)
            if ((g.cinv >> s) & 1) d ^= col[g.lb[s]];
        for (int i = 0; i < g.k; ++i) col[g.lb[i]] = nc[i];
    }

    // bank-select bits of a byte offset for the access width in use
    static uint32_t bank_mask(bool wide) { return wide ? 0x70u : 0x78u; }

    bool vec_possible(const int* R, int nr) const {
        if (!C64) return false;
        bool has0 = false;
        for (int i = 0; i < nr; ++i)
            if (R[i] == 0) has0 = true;
        if (!has0 || col[0] != 8u || (d & 8u)) return false;
        for (int t = 1; t < T; ++t)
            if (col[t] & 8u) return false;
        return true;
    }

    int emit_round(std::vector<int>& R, const std::vector<int>& members) {
        if (q->nrounds >= LP_MAX_ROUNDS) return fail(TCB200_ERR_CAPACITY, "gate pass needs more than %d rounds", LP_MAX_ROUNDS);
        // ---- filler bits ----
        bool inR[LP_MAX_T] = {false};
        for (int b : R) inR[b] = true;
        if ((int)R.size() < LP_RB && C64 && !inR[0]) {
            // logical bit 0 as a filler makes 16-byte accesses possible when its column is pure
            std::vector<int> R2 = R;
            R2.push_back(0);
            if (vec_possible(R2.data(), (int)R2.size())) {
                R.push_back(0);
                inR[0] = true;
            }
        }
        while ((int)R.size() < LP_RB) {
            const bool wide = !C64 || vec_possible(R.data(), (int)R.size());
            const uint32_t mask = bank_mask(wide);
            int best = -1, best_rank = -1;
            for (int t = T - 1; t >= 0; --t) {
                if (inR[t]) continue;
                uint32_t gc[LP_MAX_T];
                int ng = 0;
                for (int u = 0; u < T; ++u)
                    if (!inR[u] && u != t) gc[ng++] = col[u];
                const int rk = gf2_rank(gc, ng, mask);
                if (rk > best_rank) {
                    best_rank = rk;
                    best = t;
                }
            }
            R.push_back(best);
            inR[best] = true;
        }
        std::sort(R.begin(), R.end());
        const bool vec = vec_possible(R.data(), LP_RB);
        const bool wide = !C64 || vec;
        LRound& r = q->r[q->nrounds++];
        memset(&r, 0, sizeof(r));
        r.d = d;
        for (int p = 0; p < LP_RB; ++p) r.rcol[p] = col[R[p]];
        // ---- lane order of the group bits: independent bank columns first ----
        const int ngb = T - LP_RB;
        int gl[LP_MAX_T];
        int ng = 0;
        for (int t = 0; t < T; ++t)
            if (!inR[t]) gl[ng++] = t;
        const uint32_t mask = bank_mask(wide);
        const int need = wide ? 3 : 4;
        uint32_t chosen[LP_MAX_T];
        int order[LP_MAX_T];
        bool used[LP_MAX_T] = {false};
        int no = 0;
        for (int i = 0; i < ng && no < need; ++i) {
            chosen[no] = col[gl[i]];
            if (gf2_rank(chosen, no + 1, mask) == no + 1) {
                order[no++] = gl[i];
                used[i] = true;
            }
        }
        if (no < need && ngb >= need) info.conflicts++;
        for (int i = 0; i < ng; ++i)
            if (!used[i]) order[no++] = gl[i];
        for (int i = 0; i < ngb; ++i) r.gcol[i] = col[order[i]];
        if (vec) info.vec_rounds++;
        // ---- micro-ops ----
        int nc = 0;
        int last_dg = -1;  // matrix offset of a diagonal table the next diagonal op can merge into
        const int nb = blob ? batch : 1;
        std::vector<cd> dgtab((size_t)nb * 16);
        auto flush_dg = [&]() {  // the finished table of a run of diagonal ops goes to the parameter bank / blob
            if (last_dg < 0) return;
            for (int b = 0; b < nb; ++b)
                for (int x = 0; x < 16; ++x) put(last_dg + x, b, dgtab[(size_t)b * 16 + x]);
            last_dg = -1;
        };
        for (int oi : members) {
            const GOp& g = ops[oi];
            int pos[4];
            for (int j = 0; j < g.k; ++j) pos[j] = (int)(std::find(R.begin(), R.end(), g.lb[j]) - R.begin());
            if (g.kind == GK_DIAG) {
                const int D = 1 << g.k;
                int xidx[16];  // table index -> entry of this op's diagonal
                for (int x = 0; x < 16; ++x) {
                    int idx = 0;
                    for (int j = 0; j < g.k; ++j) idx |= ((x >> pos[j]) & 1) << j;
                    xidx[x] = idx;
                }
                const size_t bstep = g.bm > 1 ? (size_t)D : 0;
                if (last_dg >= 0) {  // consecutive diagonal ops: one table (product kept in double, written once)
                    for (int b = 0; b < nb; ++b) {
                        const cd* e = g.m.data() + (size_t)b * bstep;
                        cd* t = dgtab.data() + (size_t)b * 16;
                        for (int x = 0; x < 16; ++x) t[x] *= e[xidx[x]];
                    }
                    continue;
                }
                if (nmat + 16 > LP_MAT_ELEMS) return fail(TCB200_ERR_CAPACITY, "gate pass matrices exceed the parameter bank");
                if (nc >= LP_MAX_CODES) return fail(TCB200_ERR_CAPACITY, "more than %d micro-ops in a round", LP_MAX_CODES);
                for (int b = 0; b < nb; ++b) {
                    const cd* e = g.m.data() + (size_t)b * bstep;
                    cd* t = dgtab.data() + (size_t)b * 16;
                    for (int x = 0; x < 16; ++x) t[x] = e[xidx[x]];
                }
                r.code[nc++] = LOP_DG | ((uint32_t)nmat << 8);
                last_dg = nmat;
                nmat += 16;
                info.fma_per_amp += 4;
                continue;
            }
            flush_dg();
            // dense: sort the positions ascending and permute the matrix index bits to match
            const int k = g.k, D = 1 << k;
            int srt[4];
            for (int j = 0; j < k; ++j) srt[j] = j;
            std::sort(srt, srt + k, [&](int a, int b) { return pos[a] < pos[b]; });
            auto remap = [&](int i) {  // index in sorted-position order -> index in the op's own order
                int o = 0;
                for (int s = 0; s < k; ++s)
                    if ((i >> s) & 1) o |= 1 << srt[s];
                return o;
            };
            if (nmat + D * D > LP_MAT_ELEMS) return fail(TCB200_ERR_CAPACITY, "gate pass matrices exceed the parameter bank");
            if (nc >= LP_MAX_CODES) return fail(TCB200_ERR_CAPACITY, "more than %d micro-ops in a round", LP_MAX_CODES);
            int rm[8], smap[64];
            for (int i = 0; i < D; ++i) rm[i] = remap(i);
            for (int i = 0; i < D; ++i)
                for (int j = 0; j < D; ++j) smap[i * D + j] = 2 * (rm[i] * D + rm[j]);
            for (int b = 0; b < nb; ++b) {
                const double* gm = g.src + (size_t)(g.bm > 1 ? b : 0) * 2 * D * D;
                for (int e = 0; e < D * D; ++e) put(nmat + e, b, cd(gm[smap[e]], gm[smap[e] + 1]));
            }
            uint32_t opc = 0;
            const int p0 = pos[srt[0]], p1 = k > 1 ? pos[srt[1]] : 0, p2 = k > 2 ? pos[srt[2]] : 0;
            if (k == 1 && g.half) opc = (g.half == 1 ? LOP_RA : LOP_RB) + (uint32_t)p0;
            else if (k == 1) opc = LOP_G1 + (uint32_t)p0;
            else if (k == 2) {
                static const int pair_index[4][4] = {{-1, 0, 1, 2}, {-1, -1, 3, 4}, {-1, -1, -1, 5}, {-1, -1, -1, -1}};
                opc = LOP_G2 + (uint32_t)pair_index[p0][p1];
            } else {
                const int e = 6 - p0 - p1 - p2;  // the position left out
                opc = LOP_G3 + (uint32_t)e;
            }
            r.code[nc++] = opc | ((uint32_t)nmat << 8);
            nmat += D * D;
            info.fma_per_amp += (k == 1 && g.half) ? 4.0 : 4.0 * D;
        }
        flush_dg();
        r.ncodes = (uint32_t)nc | (vec ? 0x100u : 0u);
        info.rounds++;
        return 0;
    }

    int run() {
        const int n = (int)ops.size();
        int remaining = n;
        while (remaining > 0) {
            // index-map ops that are ready cost nothing: apply them all
            bool again = true;
            while (again) {
                again = false;
                for (int i = 0; i < n; ++i)
                    if (ops[i].kind == GK_LIN && ready(i)) {
                        apply_lin(ops[i]);
                        done[i] = 1;
                        --remaining;
                        info.nlin++;
                        again = true;
                    }
            }
            if (remaining == 0) break;
            // open a round with the first ready op, then add ready ops while <= LP_RB bits are touched
            std::vector<int> R, members;
            int budget = LP_MAX_CODES;
            for (;;) {
                int best = -1, best_new = 99;
                for (int i = 0; i < n; ++i) {
                    if (ops[i].kind == GK_LIN || !ready(i)) continue;
                    int nnew = 0;
                    for (int j = 0; j < ops[i].k; ++j)
                        if (std::find(R.begin(), R.end(), ops[i].lb[j]) == R.end()) ++nnew;
                    if ((int)R.size() + nnew > LP_RB) continue;
                    if (nnew < best_new) {
                        best_new = nnew;
                        best = i;
                        if (nnew == 0) break;
                    }
                }
                if (best < 0 || budget == 0) break;
                for (int j = 0; j < ops[best].k; ++j)
                    if (std::find(R.begin(), R.end(), ops[best].lb[j]) == R.end()) R.push_back(ops[best].lb[j]);
                members.push_back(best);
                done[best] = 1;
                --remaining;
                --budget;
                if (ops[best].kind == GK_DIAG) info.ndiag++;
                else info.ndense++;
            }
            if (members.empty()) return fail(TCB200_ERR_ARG, "gate pass scheduler made no progress");
            const int rc = emit_round(R, members);
            if (rc) return rc;
        }
        q->out.d = d;
        for (int t = 0; t < T; ++t) q->out.col[t] = col[t];
        bool vec = C64 && col[0] == 8u && !(d & 8u);
        for (int t = 1; t < T && vec; ++t)
            if (col[t] & 8u) vec = false;
        q->out.vec = vec ? 1u : 0u;
        info.mat_elems = nmat;
        return 0;
    }
};

// parameter block of one gate pass; 0 on success
template <typename Real>
int fill_lpass(LPassParams<Real>& q, LPassInfo& info, void* state, int nbits, int nops, const int* ops_k, const int* ops_bits,
               const double* mats, int n_hi, const int* tile_hi, const int* ops_batched = nullptr, int batch = 1,
               ME<Real>* blob = nullptr, int swz_mode = SWZ_SW, bool tight_stride = false) {
    using C = typename CT<Real>::type;
    q.state = static_cast<C*>(state);
    q.bmats = nullptr;
    q.bstride = 0;
    const int tile_bits = pass_tile_bits(sizeof(Real) == 4 ? TCB200_C64 : TCB200_C128);
    if (tile_bits > LP_MAX_T) return fail(TCB200_ERR_UNSUPPORTED, "pass tile of 2^%d amplitudes exceeds the gate-pass limit 2^%d", tile_bits, LP_MAX_T);
    // 9 gathered bits (128-byte rows of complex64) only in the production shape, whose staging does
    // not go through the 256-entry shared-memory row table
    const bool fast_shape = nbits > tile_bits && tile_bits - (sizeof(C) == 8 ? 1 : 0) == 12;
    int rc = make_geom_hi(nbits, tile_bits, nbits <= tile_bits ? 0 : n_hi, tile_hi, &q.g, fast_shape ? 9 : 8);
    if (rc) return rc;
    if (q.g.T < LP_RB) return fail(TCB200_ERR_UNSUPPORTED, "gate pass needs a state of at least %d bits", LP_RB);
    Scheduler<Real> s;
    s.T = q.g.T;
    s.q = &q;
    memset(&s.info, 0, sizeof(s.info));
    s.swz_mode = swz_mode;
    {
        const char* e = getenv("TCB200_LAYOUT_HW128");  // planning statistics (dry runs only)
        if (e && e[0] == '1' && !state) s.swz_mode = SWZ_HW128;
    }
    s.init_layout();
    s.batch = batch;
    s.blob = blob;
    const int* b = ops_bits;
    const double* mp = mats;
    for (int o = 0; o < nops; ++o) {
        const int k = ops_k[o];
        if (k < 1 || k > 4) return fail(TCB200_ERR_UNSUPPORTED, "gate of %d bits inside a gate pass", k);
        int lb[4];
        for (int i = 0; i < k; ++i) {
            if (b[i] < 0 || b[i] >= nbits) return fail(TCB200_ERR_ARG, "bit %d out of range", b[i]);
            if (i > 0 && b[i] <= b[i - 1]) return fail(TCB200_ERR_ARG, "bits must be strictly ascending");
            lb[i] = local_bit(q.g, b[i]);
            if (lb[i] < 0) return fail(TCB200_ERR_ARG, "bit %d is not inside the tile", b[i]);
        }
        const int bm = (ops_batched && ops_batched[o]) ? batch : 1;
        rc = classify(k, lb, mp, bm, s.ops);
        if (rc) return rc;
        b += k;
        mp += (size_t)bm * (2ll << (2 * k));
    }
    if (tight_stride && blob) {
        // per-element blob stride = what the classified ops can need at most (diagonal merges only lower it):
        // the chunks that travel to the constant bank are then contiguous, no compaction copy
        size_t ub = 0;
        for (const GOp& g : s.ops) ub += g.kind == GK_DENSE ? (size_t)1 << (2 * g.k) : (g.kind == GK_DIAG ? 16 : 0);
        s.bstride = std::min<size_t>(LP_MAT_ELEMS, std::max<size_t>(ub, 1));
    }
    q.bstride = (int)s.bstride;
    q.nrounds = 0;
    q.ngb = q.g.T - LP_RB;
    q.tb = q.ngb < 8 ? q.ngb : 8;
    // staging threads: one 16-byte unit per lane per iteration, at least a warp
    const int units_log = q.g.T - (sizeof(C) == 8 ? 1 : 0);
    q.stb = q.tb;
    if (q.stb < 5) q.stb = 5;
    if (q.stb > units_log && units_log >= 5) q.stb = units_log;
    s.build_deps();
    rc = s.run();
    if (rc) return rc;
    info = s.info;
    // ---- staging constants of the production shape ----
    constexpr int APU = 16 / (int)sizeof(C);
    constexpr int SH = APU == 2 ? 1 : 0;
    q.fast = (q.stb == 8 && units_log == 12) ? 1 : 0;
    if (q.fast) {
        auto goff_of = [&](uint32_t e) {  // amplitude offset of tile-local element e
            return row_offset(q.g, e >> q.g.lrow) + (uint64_t)(e & ((1u << q.g.lrow) - 1u));
        };
        for (int b = 0; b < 9; ++b) q.stage.bit_off[b] = goff_of(1u << b);
        for (int i = 0; i < LP_FAST_ITERS; ++i) {
            const uint32_t u = (uint32_t)i << 8;
            q.stage.goff[i] = goff_of(u * APU);
            q.stage.sin[i] = swz_unit(u) << 4;
            uint32_t x = 0;
            for (int t = 0; t < 4; ++t)
                if ((i >> t) & 1) x ^= q.out.col[8 + t + SH];
            q.stage.sout[i] = x;
        }
    }
    return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------
// BATCHED (vmap): batch element blockIdx.y reads its own matrices from global memory
// (p.bmats + blockIdx.y * p.bstride; warp-uniform, L1-resident loads) instead of the parameter bank
template <typename Real, bool BATCHED>
__global__ void __launch_bounds__(256, (sizeof(Real) == 4 && !BATCHED) ? 3 : 2) lpass_kernel(const __grid_constant__ LPassParams<Real> p) {
    using C = typename CT<Real>::type;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    C* tile = reinterpret_cast<C*>(smem_raw);
    __shared__ uint64_t rowoff[256];

    const int tid = threadIdx.x;
    const int nthr = blockDim.x;
    if (tid < (1 << p.g.h)) rowoff[tid] = row_offset(p.g, tid);
    C* vec = p.state + ((uint64_t)blockIdx.y << p.g.n);
    const uint64_t base = tile_base(p.g, blockIdx.x);
    __syncthreads();
    stage_in<C, SWZ_SW>(p.g, vec, base, tile, rowoff, tid, nthr);
    cp_async_wait_all();
    __syncthreads();
    const bool worker = (tid >> p.tb) == 0;
    for (int r = 0; r < p.nrounds; ++r) {
        if (worker) lround_thread<C, Real>(smem_raw, p.r[r], BATCHED ? p.bmats + (size_t)blockIdx.y * p.bstride : p.m, (uint32_t)tid, p.tb, p.ngb);
        __syncthreads();
    }
    lstage_out_thread<C>(p.g, vec, base, smem_raw, rowoff, p.out, tid, nthr, p.stb);
}

// production shape (LPassParams::fast): 256 threads, one 64 KiB tile, host-precomputed staging
// constants, compile-time loop structure -- no shared-memory row table, no per-unit index math
template <typename Real, bool BATCHED>
__global__ void __launch_bounds__(256, (sizeof(Real) == 4 && !BATCHED) ? 3 : 2) lpass_fast_kernel(const __grid_constant__ LPassParams<Real> p) {
    using C = typename CT<Real>::type;
    constexpr int NIT = sizeof(C) == 8 ? 2 : 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t tid = threadIdx.x;
    C* vec = p.state + ((uint64_t)blockIdx.y << p.g.n) + tile_base(p.g, blockIdx.x);
    lstage_in_fast<C>(p.stage, vec, smem_raw, tid);
    cp_async_wait_all();
    __syncthreads();
    for (int r = 0; r < p.nrounds; ++r) {
        lround_thread_fast<C, Real, NIT>(smem_raw, p.r[r], BATCHED ? p.bmats + (size_t)blockIdx.y * p.bstride : p.m, tid);
        __syncthreads();
    }
    lstage_out_fast<C>(p.stage, p.out, vec, smem_raw, tid);
}

// ------------------------------------------------------------------------------------------------
// pipelined production shape (complex64): one persistent CTA per SM, three 64 KiB tile buffers
// ------------------------------------------------------------------------------------------------
// lpass_fast_kernel runs load -> rounds -> write-back serially inside a CTA and relies on the three
// CTAs of an SM to overlap each other's phases; measured, a pass costs about 0.66 x (HBM time) more
// than max(HBM, FP32).  Here the phases are split by warp role instead:
//
//   4 producer warps                           2 compute groups of 8 warps (tiles j = g, g + 2, ...)
//   ----------------------------------         -------------------------------------------------
//   wait empty[j % 3]                          wait full[j % 3]
//   LDGSTS tile j -> buf[j % 3]                rounds on buf[j % 3]   (bar.sync 1 + g between rounds)
//   cp.async.mbarrier.arrive full[j % 3]       write-back through the final index map
//                                              arrive empty[j % 3]
//
// so while the two groups work on two buffers the third one is always being filled, and the two
// groups are in different phases of their tiles (one in its FMA rounds while the other writes back).
// Same staging swizzle, same rounds, same write-back as lpass_fast_kernel: the parameter block is
// the same one.
constexpr int PL_STAGES = 3;
constexpr int PL_GROUP = 256;
constexpr int PL_NGROUPS = 2;
constexpr int PL_PRODUCERS = 128;
constexpr int PL_THREADS = PL_GROUP * PL_NGROUPS + PL_PRODUCERS;
constexpr uint32_t PL_TILE_BYTES = 64 * 1024;

__device__ __forceinline__ uint32_t pl_smem(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void pl_mbar_init(uint64_t* b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(pl_smem(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void pl_mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(pl_smem(b)) : "memory");
}
// arrive once every cp.async this thread has issued so far has landed (counted in the init count)
__device__ __forceinline__ void pl_cp_async_arrive(uint64_t* b) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(pl_smem(b)) : "memory");
}
__device__ __forceinline__ void pl_mbar_wait(uint64_t* b, uint32_t parity) {
    const uint32_t a = pl_smem(b);
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
__device__ __forceinline__ void pl_bar_group(uint32_t id) { asm volatile("bar.sync %0, %1;\n" ::"r"(id), "n"(PL_GROUP) : "memory"); }

template <typename Real>
__global__ void __launch_bounds__(PL_THREADS, 1) lpass_pipe_kernel(const __grid_constant__ LPassParams<Real> p) {
    using C = typename CT<Real>::type;
    constexpr int APU = 16 / (int)sizeof(C);
    constexpr int NIT = sizeof(C) == 8 ? 2 : 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full[PL_STAGES];
    __shared__ __align__(8) uint64_t empty[PL_STAGES];

    const uint32_t tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < PL_STAGES; ++s) {
            pl_mbar_init(&full[s], PL_PRODUCERS);
            pl_mbar_init(&empty[s], PL_GROUP);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    const uint64_t first = blockIdx.x, stride = gridDim.x;
    const uint32_t nmine = p.ntiles > first ? (uint32_t)((p.ntiles - first + stride - 1) / stride) : 0u;
    const uint64_t vmask = (1ull << p.vtl) - 1ull;
    auto tile_ptr = [&](uint32_t j) -> C* {
        const uint64_t t = first + (uint64_t)j * stride;
        return p.state + ((t >> p.vtl) << p.g.n) + tile_base(p.g, t & vmask);
    };

    // warp-uniform role (the shuffle tells the compiler so: without it everything below counts as
    // divergent code and the matrices leave the uniform registers -- LDC + MOV instead of LDCU / UR operands)
    const uint32_t warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    if (warp >= PL_GROUP * PL_NGROUPS / 32) {
        // ---- producers: thread pt stages the units of the virtual staging threads pt and pt + 128 ----
        const uint32_t pt = tid - PL_GROUP * PL_NGROUPS;
        const uint64_t go0 = lstage_thread_goff<APU>(p.stage, pt), go1 = lstage_thread_goff<APU>(p.stage, pt + PL_PRODUCERS);
        const uint32_t s0 = swz_unit(pt) << 4, s1 = swz_unit(pt + PL_PRODUCERS) << 4;
        uint32_t s = 0, ph = 1;  // ph: parity of the previous completion of empty[s] (none yet for the first lap)
        for (uint32_t j = 0; j < nmine; ++j) {
            if (j >= PL_STAGES) pl_mbar_wait(&empty[s], ph);
            unsigned char* buf = smem_raw + s * PL_TILE_BYTES;
            const C* v = tile_ptr(j);
            const C* v0 = v + go0;
            const C* v1 = v + go1;
#pragma unroll
            for (int i = 0; i < LP_FAST_ITERS; ++i) cp_async16(buf + (s0 ^ p.stage.sin[i]), v0 + p.stage.goff[i]);
#pragma unroll
            for (int i = 0; i < LP_FAST_ITERS; ++i) cp_async16(buf + (s1 ^ p.stage.sin[i]), v1 + p.stage.goff[i]);
            pl_cp_async_arrive(&full[s]);
            if (++s == PL_STAGES) {
                s = 0;
                ph ^= 1u;
            }
        }
        cp_async_wait_all();
        return;
    }

    // ---- compute groups ----
    const uint32_t grp = warp >> 3, lt = tid & (PL_GROUP - 1);
    uint32_t s = grp, ph = 0;  // tile j lives in buffer j % 3; ph = (j / 3) & 1
    for (uint32_t j = grp; j < nmine; j += PL_NGROUPS) {
        unsigned char* buf = smem_raw + s * PL_TILE_BYTES;
        pl_mbar_wait(&full[s], ph);
        for (int r = 0; r < p.nrounds; ++r) {
            lround_thread_fast<C, Real, NIT>(buf, p.r[r], p.m, lt);
            pl_bar_group(1 + grp);
        }
        lstage_out_fast<C>(p.stage, p.out, tile_ptr(j), buf, lt);
        pl_mbar_arrive(&empty[s]);
        s += PL_NGROUPS;
        if (s >= PL_STAGES) {
            s -= PL_STAGES;
            ph ^= 1u;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// TMA-staged production shape: the tile arrives as cp.async.bulk.tensor boxes issued by ONE thread
// ------------------------------------------------------------------------------------------------
// Same rounds and write-back as lpass_fast_kernel; the 16 LDGSTS + address arithmetic per thread of
// the stage-in (10 % of the issued instructions of a pass, and its share of the MIO queue) are
// replaced by <= 64 bulk tensor copies (one per thread of the first two warps) that complete on an mbarrier.  The tile then
// sits in the hardware SWIZZLE_128B layout, which is one more affine index map for the scheduler
// (Scheduler::swz_mode = SWZ_HW128).
__device__ __forceinline__ void pl_mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(pl_smem(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pl_tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n"
        ::"r"(pl_smem(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(pl_smem(bar)), "r"(0), "r"(c1), "r"(0), "r"(0), "r"(0)
        : "memory");
}

template <typename Real>
__global__ void __launch_bounds__(256, 3) lpass_tma_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ LPassParams<Real> p) {
    using C = typename CT<Real>::type;
    constexpr int NIT = sizeof(C) == 8 ? 2 : 1;
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full;
    // SWIZZLE_128B atoms are 1024 bytes: align the tile by hand (the launch reserves the slack)
    unsigned char* tile = smem_raw + ((1024u - (pl_smem(smem_raw) & 1023u)) & 1023u);
    const uint32_t tid = threadIdx.x;
    const uint64_t base = ((uint64_t)blockIdx.y << p.g.n) + tile_base(p.g, blockIdx.x);
    if (tid == 0) {
        pl_mbar_init(&full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        pl_mbar_expect_tx(&full, PL_TILE_BYTES);
    }
    __syncthreads();  // the barrier is armed before any copy can complete on it and before anyone polls it
    // one box per thread: a single issuing thread needs ~4 us for the 64 boxes of a 9-gathered-bit tile
    if (tid < (1u << p.tp.nextra)) {
        const uint32_t box_bytes = (uint32_t)sizeof(C) << p.tp.box_amps_log2;
        pl_tma_load_5d(tile + tid * box_bytes, &tmap, &full, (int)((base + tma_box_offset(p.tp, tid)) >> p.tp.line_bits));
    }
    pl_mbar_wait(&full, 0);
    for (int r = 0; r < p.nrounds; ++r) {
        lround_thread_fast<C, Real, NIT>(tile, p.r[r], p.m, tid);
        __syncthreads();
    }
    lstage_out_fast<C>(p.stage, p.out, p.state + base, tile, tid);
}

// TCB200_GATE_TMA=1: stage the tiles of the production shape by TMA (complex64, shared matrices)
static bool gate_tma_enabled() {
    const char* e = getenv("TCB200_GATE_TMA");
    return e && e[0] == '1';
}

// Variant B of the pipeline: two compute groups of 512 threads (one 16-amplitude group per thread and
// round), no producer warps -- the group that has written a tile back refills the buffer itself (for
// the tile the OTHER group will work on three tiles later).  32 warps at 64 registers.
constexpr int PB_GROUP = 512;
__device__ __forceinline__ void pb_bar_group(uint32_t id) { asm volatile("bar.sync %0, %1;\n" ::"r"(id), "n"(PB_GROUP) : "memory"); }

template <typename C, typename Real>
__device__ __forceinline__ void lround_thread_512(unsigned char* tile, const LRound& R, const ME<Real>* mats, uint32_t tid) {
    uint32_t b = R.d;
#pragma unroll
    for (int i = 0; i < 9; ++i) b ^= (0u - ((tid >> i) & 1u)) & R.gcol[i];
    const bool vec = sizeof(C) == 8 && ((R.ncodes >> 8) & 1u);
    lround_group<C, Real>(tile, b, R, mats, vec);
}

template <typename Real>
__global__ void __launch_bounds__(2 * PB_GROUP, 1) lpass_pipe2_kernel(const __grid_constant__ LPassParams<Real> p) {
    using C = typename CT<Real>::type;
    constexpr int APU = 16 / (int)sizeof(C);
    constexpr int SH = APU == 2 ? 1 : 0;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full[PL_STAGES];

    const uint32_t tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < PL_STAGES; ++s) pl_mbar_init(&full[s], PB_GROUP);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    const uint64_t first = blockIdx.x, stride = gridDim.x;
    const uint32_t nmine = p.ntiles > first ? (uint32_t)((p.ntiles - first + stride - 1) / stride) : 0u;
    const uint64_t vmask = (1ull << p.vtl) - 1ull;
    auto tile_ptr = [&](uint32_t j) -> C* {
        const uint64_t t = first + (uint64_t)j * stride;
        return p.state + ((t >> p.vtl) << p.g.n) + tile_base(p.g, t & vmask);
    };
    const uint32_t grp = __shfl_sync(0xffffffffu, tid >> 9, 0), lt = tid & (PB_GROUP - 1);  // warp-uniform
    // staging: thread lt owns the units (lt & 255) + 256 i for i in [8 * (lt >> 8), 8 * (lt >> 8) + 8)
    const uint32_t vt = lt & 255u, half = lt >> 8;
    const uint64_t tg = lstage_thread_goff<APU>(p.stage, vt);
    const uint32_t sin0 = swz_unit(vt) << 4;
    uint32_t aout = p.out.d;
#pragma unroll
    for (int i = 0; i < 8; ++i) aout ^= (0u - ((vt >> i) & 1u)) & p.out.col[i + SH];
    const bool ovec = APU == 1 || p.out.vec;
    const uint32_t oc0 = p.out.col[0];

    auto issue_load = [&](uint32_t j, uint32_t s) {
        unsigned char* buf = smem_raw + s * PL_TILE_BYTES;
        const C* v = tile_ptr(j) + tg;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int ii = (int)half * 8 + i;
            cp_async16(buf + (sin0 ^ p.stage.sin[ii]), v + p.stage.goff[ii]);
        }
        pl_cp_async_arrive(&full[s]);
    };
    // prologue: group 0 fills buffers 0 and 2 (its first two tiles), group 1 buffer 1
    if (grp < nmine) issue_load(grp, grp);
    if (grp == 0 && 2 < nmine) issue_load(2, 2);

    uint32_t s = grp, ph = 0;
    for (uint32_t j = grp; j < nmine; j += 2) {
        unsigned char* buf = smem_raw + s * PL_TILE_BYTES;
        pl_mbar_wait(&full[s], ph);
        for (int r = 0; r < p.nrounds; ++r) {
            lround_thread_512<C, Real>(buf, p.r[r], p.m, lt);
            pb_bar_group(1 + grp);
        }
        {
            C* g = tile_ptr(j) + tg;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int ii = (int)half * 8 + i;
                const uint32_t ai = aout ^ p.stage.sout[ii];
                Unit16 q;
                if (ovec) {
                    q = *reinterpret_cast<const Unit16*>(buf + ai);
                } else {
                    C* qc = reinterpret_cast<C*>(&q);
                    qc[0] = *reinterpret_cast<const C*>(buf + ai);
                    qc[1 % APU] = *reinterpret_cast<const C*>(buf + (ai ^ oc0));
                }
                *reinterpret_cast<Unit16*>(g + p.stage.goff[ii]) = q;
            }
        }
        if (j + PL_STAGES < nmine) {
            pb_bar_group(1 + grp);  // every thread of the group has read its part of the buffer
            issue_load(j + PL_STAGES, s);
        }
        s += 2;
        if (s >= PL_STAGES) {
            s -= PL_STAGES;
            ph ^= 1u;
        }
    }
    cp_async_wait_all();
}

static void launch_tma(const CUtensorMap& tmap, const LPassParams<float>& q, dim3 grid, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(lpass_tma_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(PL_TILE_BYTES + 1024));
        attr = true;
    }
    lpass_tma_kernel<float><<<grid, 256, PL_TILE_BYTES + 1024, st>>>(tmap, q);
}
static void launch_tma(const CUtensorMap&, const LPassParams<double>&, dim3, cudaStream_t) {}

static void launch_pipe(const LPassParams<float>& q, unsigned grid, cudaStream_t st) {
    const char* e = getenv("TCB200_PIPE");
    if (e && e[0] == '2') lpass_pipe2_kernel<float><<<grid, 2 * PB_GROUP, PL_STAGES * PL_TILE_BYTES, st>>>(q);
    else lpass_pipe_kernel<float><<<grid, PL_THREADS, PL_STAGES * PL_TILE_BYTES, st>>>(q);
}
static void launch_pipe(const LPassParams<double>&, unsigned, cudaStream_t) {}

// Opt-in (TCB200_PIPE=1: producer warps + two groups of 8 warps; =2: two self-loading groups of 16
// warps at 64 registers).  Both are bit-identical to lpass_fast_kernel and both are SLOWER on the
// measured circuits (config-4 recipe at n = 30: 127 ms fast, 143 ms / 185 ms pipelined;
// profiles/README.md "Pipelined gate pass"): the rounds are issue-bound, and 16 always-computing warps
// issue less than the 24 warps of three independent CTAs even though those stall on their own staging.
// (read at every launch: the tests and the A/B scripts flip it inside one process)
static bool pipe_enabled() {
    const char* e = getenv("TCB200_PIPE");
    return e && (e[0] == '1' || e[0] == '2');
}
// test knobs: TCB200_PIPE_MIN_TILES (tiles from which the pipeline is used; default 4 per SM),
// TCB200_PIPE_GRID (persistent CTAs; default one per SM)
static long pipe_knob(const char* name, long dflt) {
    const char* e = getenv(name);
    if (!e || !e[0]) return dflt;
    const long v = atol(e);
    return v > 0 ? v : dflt;
}

static int lp_sm_count() {
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

// BATCHED through the constant bank: the per-element matrices of a CHUNK of batch elements sit in a
// __constant__ array (filled by cudaMemcpyToSymbolAsync before the launch), element blockIdx.y at
// blockIdx.y * p.bstride.  The offset is warp-uniform, so the matrices reach the FMAs exactly as in the
// unbatched kernel -- LDCU -> uniform registers, 80 registers, 3 CTAs / SM -- instead of by per-thread
// L1 loads into vector registers (128 registers, 2 CTAs / SM: 8.8 ms against 2.5 ms per pass of 2^30
// amplitudes, profiles/r2_config3_launches_summary.txt).
constexpr int LP_CBANK_ELEMS = 3840;  // 60 KiB of the 64 KiB constant bank
__constant__ uint4 g_cmat[LP_CBANK_ELEMS];

template <typename Real>
__global__ void __launch_bounds__(256, sizeof(Real) == 4 ? 3 : 2) lpass_fast_cbank_kernel(const __grid_constant__ LPassParams<Real> p) {
    using C = typename CT<Real>::type;
    constexpr int NIT = sizeof(C) == 8 ? 2 : 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t tid = threadIdx.x;
    C* vec = p.state + ((uint64_t)blockIdx.y << p.g.n) + tile_base(p.g, blockIdx.x);
    const ME<Real>* mats = reinterpret_cast<const ME<Real>*>(g_cmat) + blockIdx.y * p.bstride;
    lstage_in_fast<C>(p.stage, vec, smem_raw, tid);
    cp_async_wait_all();
    __syncthreads();
    for (int r = 0; r < p.nrounds; ++r) {
        lround_thread_fast<C, Real, NIT>(smem_raw, p.r[r], mats, tid);
        __syncthreads();
    }
    lstage_out_fast<C>(p.stage, p.out, vec, smem_raw, tid);
}

// TCB200_CBANK=0: per-element matrices from global memory (lpass_fast_kernel<Real, true>)
static bool cbank_enabled() {
    const char* e = getenv("TCB200_CBANK");
    return !(e && e[0] == '0');
}

// TCB200_TIMING=1: host microseconds spent in (0) classification + scheduling + matrix packing and (1) everything
// after it (copies, launches, waits for a pinned buffer), printed when the library is unloaded
struct HostTimerTotals {
    double us[2] = {0, 0};
    long calls = 0;
    bool on = false;
    HostTimerTotals() {
        const char* e = getenv("TCB200_TIMING");
        on = e && e[0] == '1';
    }
    ~HostTimerTotals() {
        if (on) fprintf(stderr, "[tcb200 timing] gate passes: %ld calls, plan+pack %.1f ms, copy+launch %.1f ms\n", calls, us[0] / 1e3, us[1] / 1e3);
    }
};
static HostTimerTotals g_host_timer;
struct HostTimer {
    int slot;
    bool running;
    std::chrono::steady_clock::time_point t0;
    explicit HostTimer(int s) : slot(s), running(g_host_timer.on), t0(std::chrono::steady_clock::now()) {}
    void stop() {
        if (!running) return;
        running = false;
        g_host_timer.us[slot] += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
        if (slot == 0) g_host_timer.calls++;
    }
    ~HostTimer() { stop(); }
};

// pinned staging + events for the per-element matrix blobs of batched passes (two buffers, so that
// the host can fill the blob of pass i + 1 while the copy of pass i is still in flight)
struct BlobStage {
    void* host[2] = {nullptr, nullptr};
    size_t cap[2] = {0, 0};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    int next = 0;
};

template <typename Real>
static int launch_lpass(void* state, int nbits, int nops, const int* ops_k, const int* ops_bits, const double* mats, int n_hi,
                        const int* tile_hi, int64_t batch, cudaStream_t st, LPassInfo* info_out, const int* ops_batched = nullptr,
                        void* workspace = nullptr, size_t ws_bytes = 0) {
    using C = typename CT<Real>::type;
    static thread_local LPassParams<Real>* tp = nullptr;
    if (!tp) tp = new LPassParams<Real>();
    LPassParams<Real>& q = *tp;
    LPassInfo info;
    bool batched = false;
    if (ops_batched)
        for (int o = 0; o < nops; ++o) batched = batched || ops_batched[o] != 0;
    ME<Real>* blob = nullptr;
    static thread_local BlobStage bs;
    int slot = 0;
    const size_t blob_bytes = (size_t)batch * LP_MAT_ELEMS * sizeof(ME<Real>);
    if (batched && state) {
        if (batch < 1 || batch > 65535) return fail(TCB200_ERR_ARG, "batch=%lld out of range", (long long)batch);
        if (!workspace || ws_bytes < blob_bytes) return fail(TCB200_ERR_WORKSPACE, "batched gate pass needs %zu bytes of workspace", blob_bytes);
        slot = bs.next;
        bs.next ^= 1;
        if (bs.cap[slot] < blob_bytes) {
            if (bs.host[slot]) TCB_CUDA(cudaFreeHost(bs.host[slot]));
            TCB_CUDA(cudaMallocHost(&bs.host[slot], blob_bytes));
            bs.cap[slot] = blob_bytes;
        }
        if (!bs.ev[slot]) TCB_CUDA(cudaEventCreateWithFlags(&bs.ev[slot], cudaEventDisableTiming));
        else TCB_CUDA(cudaEventSynchronize(bs.ev[slot]));  // the copy that last used this buffer is done
        blob = static_cast<ME<Real>*>(bs.host[slot]);
    }
    std::vector<ME<Real>> dry;
    if (batched && !state) {  // dry run
        dry.resize((size_t)batch * LP_MAT_ELEMS);
        blob = dry.data();
    }
    alignas(64) CUtensorMap tmap;
    bool use_tma = state && !batched && sizeof(Real) == 4 && gate_tma_enabled();
    const bool want_cbank = batched && state && cbank_enabled();
    HostTimer timer_fill(0);
    int rc = fill_lpass<Real>(q, info, state, nbits, nops, ops_k, ops_bits, mats, n_hi, tile_hi, ops_batched, (int)batch, blob,
                              use_tma ? SWZ_HW128 : SWZ_SW, want_cbank);
    if (rc) return rc;
    if (use_tma && (!q.fast || tma_encode_state_map(state, nbits, 1, q.g, batch, &q.tp, &tmap) != 0)) {
        use_tma = false;  // not eligible (small state, alignment, coordinate range): plan again for the LDGSTS layout
        rc = fill_lpass<Real>(q, info, state, nbits, nops, ops_k, ops_bits, mats, n_hi, tile_hi, ops_batched, (int)batch, blob);
        if (rc) return rc;
    }
    timer_fill.stop();
    HostTimer timer_launch(1);
    if (info_out) *info_out = info;
    if (!state) return 0;  // dry run (tcb200_gate_pass_info)
    const uint64_t ntiles = 1ull << (nbits - q.g.T);
    if (ntiles > 0x7fffffffull) return fail(TCB200_ERR_UNSUPPORTED, "state too large for one grid");
    if (batch < 1 || batch > 65535) return fail(TCB200_ERR_ARG, "batch=%lld out of range", (long long)batch);
    size_t smem = (size_t)sizeof(C) << q.g.T;
    if (smem < 16) smem = 16;
    static bool attr = false;
    if (!attr) {
        TCB_CUDA(cudaFuncSetAttribute(lpass_kernel<Real, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
        TCB_CUDA(cudaFuncSetAttribute(lpass_fast_kernel<Real, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
        TCB_CUDA(cudaFuncSetAttribute(lpass_kernel<Real, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
        TCB_CUDA(cudaFuncSetAttribute(lpass_fast_kernel<Real, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
        attr = true;
    }
    dim3 grid((unsigned)ntiles, (unsigned)batch);
    dim3 block(1u << q.stb);
    if (batched) {
        const size_t used = (size_t)info.mat_elems * sizeof(ME<Real>);
        const size_t bstride = (size_t)q.bstride;  // elements between two batch elements in the pinned blob
        if (q.fast && used && want_cbank && bstride <= (size_t)LP_CBANK_ELEMS) {
            // chunks of batch elements through the constant bank
            static bool cattr = false;
            if (!cattr) {
                TCB_CUDA(cudaFuncSetAttribute(lpass_fast_cbank_kernel<Real>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
                cattr = true;
            }
            const int64_t per = std::max<int64_t>(1, std::min<int64_t>(batch, (int64_t)(LP_CBANK_ELEMS / bstride)));
            C* state0 = q.state;
            // the constant array is one per process: passes issued from different host threads / on different
            // streams are serialised against each other (host mutex for the issue order, an event for the device)
            static std::mutex cbank_mutex;
            static cudaEvent_t cbank_done = nullptr;
            static cudaStream_t cbank_stream = nullptr;
            std::lock_guard<std::mutex> cbank_lock(cbank_mutex);
            if (!cbank_done) TCB_CUDA(cudaEventCreateWithFlags(&cbank_done, cudaEventDisableTiming));
            else if (cbank_stream != st) TCB_CUDA(cudaStreamWaitEvent(st, cbank_done, 0));
            // the whole blob goes to the device workspace in ONE host-to-device copy; the chunks then reach the
            // constant bank by device-to-device copies (no PCIe latency between two kernels of a pass)
            const size_t total_bytes = ((size_t)(batch - 1) * bstride) * sizeof(ME<Real>) + used;
            TCB_CUDA(cudaMemcpyAsync(workspace, blob, total_bytes, cudaMemcpyHostToDevice, st));
            TCB_CUDA(cudaEventRecord(bs.ev[slot], st));
            const unsigned char* wsb = static_cast<const unsigned char*>(workspace);
            for (int64_t c0 = 0; c0 < batch; c0 += per) {
                const int64_t nb = std::min<int64_t>(per, batch - c0);
                TCB_CUDA(cudaMemcpyToSymbolAsync(g_cmat, wsb + (size_t)c0 * bstride * sizeof(ME<Real>),
                                                 ((size_t)(nb - 1) * bstride) * sizeof(ME<Real>) + used, 0, cudaMemcpyDeviceToDevice, st));
                q.state = state0 + ((uint64_t)c0 << nbits);
                lpass_fast_cbank_kernel<Real><<<dim3((unsigned)ntiles, (unsigned)nb), block, smem, st>>>(q);
                TCB_LAUNCH_CHECK("lpass_fast_cbank_kernel");
            }
            q.state = state0;
            TCB_CUDA(cudaEventRecord(cbank_done, st));
            cbank_stream = st;
            return 0;
        }
        // only the used prefix of every element's blob travels
        if (used) {
            TCB_CUDA(cudaMemcpy2DAsync(workspace, LP_MAT_ELEMS * sizeof(ME<Real>), blob, bstride * sizeof(ME<Real>), used, (size_t)batch,
                                       cudaMemcpyHostToDevice, st));
        }
        TCB_CUDA(cudaEventRecord(bs.ev[slot], st));
        q.bmats = static_cast<const ME<Real>*>(workspace);
        q.bstride = LP_MAT_ELEMS;
        if (q.fast) lpass_fast_kernel<Real, true><<<grid, block, smem, st>>>(q);
        else lpass_kernel<Real, true><<<grid, block, smem, st>>>(q);
    } else {
        const uint64_t total = ntiles * (uint64_t)batch;
        const int sms = lp_sm_count();
        if (use_tma) {
            launch_tma(tmap, q, grid, st);
            TCB_LAUNCH_CHECK("lpass_tma_kernel");
            return 0;
        }
        // the pipeline pays off once every SM walks a few tiles
        if (q.fast && sizeof(Real) == 4 && sms > 0 && pipe_enabled() && total >= (uint64_t)pipe_knob("TCB200_PIPE_MIN_TILES", 4l * sms)) {
            static bool pattr = false;
            if (!pattr) {
                TCB_CUDA(cudaFuncSetAttribute(lpass_pipe_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(PL_STAGES * PL_TILE_BYTES)));
                TCB_CUDA(cudaFuncSetAttribute(lpass_pipe2_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(PL_STAGES * PL_TILE_BYTES)));
                pattr = true;
            }
            q.vtl = nbits - q.g.T;
            q.ntiles = total;
            const uint64_t ctas = (uint64_t)pipe_knob("TCB200_PIPE_GRID", sms);
            launch_pipe(q, (unsigned)(total < ctas ? total : ctas), st);
        } else if (q.fast) lpass_fast_kernel<Real, false><<<grid, block, smem, st>>>(q);
        else lpass_kernel<Real, false><<<grid, block, smem, st>>>(q);
    }
    TCB_LAUNCH_CHECK("lpass_kernel");
    return 0;
}

#ifdef TCB200_EMU
// tests/emu only: the same parameter block and the same __host__ __device__ bodies on the CPU
template <typename Real>
static int emu_lpass(void* state, int nbits, int nops, const int* ops_k, const int* ops_bits, const double* mats, int n_hi,
                     const int* tile_hi, LPassInfo* info_out, const int* ops_batched = nullptr, int batch = 1) {
    using C = typename CT<Real>::type;
    LPassParams<Real>* q = new LPassParams<Real>();
    LPassInfo info;
    bool batched = false;
    if (ops_batched)
        for (int o = 0; o < nops; ++o) batched = batched || ops_batched[o] != 0;
    std::vector<ME<Real>> blob;
    if (batched) blob.resize((size_t)batch * LP_MAT_ELEMS);
    int rc = fill_lpass<Real>(*q, info, state, nbits, nops, ops_k, ops_bits, mats, n_hi, tile_hi, ops_batched, batch, batched ? blob.data() : nullptr);
    if (rc) {
        delete q;
        return rc;
    }
    if (info_out) *info_out = info;
    const int nthr = 1 << q->stb;
    const size_t bytes = (size_t)sizeof(C) << q->g.T;
    unsigned char* tile = static_cast<unsigned char*>(aligned_alloc(128, bytes < 128 ? 128 : bytes));
    uint64_t rowoff[256];
    for (int r = 0; r < (1 << q->g.h); ++r) rowoff[r] = row_offset(q->g, r);
    const uint64_t ntiles = 1ull << (nbits - q->g.T);
    for (int be = 0; be < batch; ++be) {
        C* vec = static_cast<C*>(state) + ((uint64_t)be << nbits);
        const ME<Real>* mm = batched ? blob.data() + (size_t)be * LP_MAT_ELEMS : q->m;
        for (uint64_t t = 0; t < ntiles; ++t) {
            const uint64_t base = tile_base(q->g, t);
            if (q->fast) {  // the bodies of lpass_fast_kernel
                constexpr int NIT = sizeof(C) == 8 ? 2 : 1;
                for (int tid = 0; tid < 256; ++tid) lstage_in_fast<C>(q->stage, vec + base, tile, (uint32_t)tid);
                for (int r = 0; r < q->nrounds; ++r)
                    for (int tid = 0; tid < 256; ++tid) lround_thread_fast<C, Real, NIT>(tile, q->r[r], mm, (uint32_t)tid);
                for (int tid = 0; tid < 256; ++tid) lstage_out_fast<C>(q->stage, q->out, vec + base, tile, (uint32_t)tid);
                continue;
            }
            for (int tid = 0; tid < nthr; ++tid) stage_in<C, SWZ_SW>(q->g, vec, base, reinterpret_cast<C*>(tile), rowoff, tid, nthr);
            for (int r = 0; r < q->nrounds; ++r)
                for (int tid = 0; tid < (1 << q->tb); ++tid) lround_thread<C, Real>(tile, q->r[r], mm, (uint32_t)tid, q->tb, q->ngb);
            for (int tid = 0; tid < nthr; ++tid) lstage_out_thread<C>(q->g, vec, base, tile, rowoff, q->out, tid, nthr, q->stb);
        }
    }
    free(tile);
    delete q;
    return 0;
}

extern "C" __attribute__((visibility("default"))) int emu_apply_gate_pass_batched(void* state, int nbits, int dtype, int nops, const int* ops_k,
                                                                                   const int* ops_bits, const double* ops_mats,
                                                                                   const int* ops_batched, int n_hi, const int* tile_hi,
                                                                                   int batch, double* info8) {
    LPassInfo info;
    memset(&info, 0, sizeof(info));
    int rc;
    if (dtype == TCB200_C64) rc = emu_lpass<float>(state, nbits, nops, ops_k, ops_bits, ops_mats, n_hi, tile_hi, &info, ops_batched, batch);
    else rc = emu_lpass<double>(state, nbits, nops, ops_k, ops_bits, ops_mats, n_hi, tile_hi, &info, ops_batched, batch);
    if (info8) {
        info8[0] = info.rounds; info8[1] = info.nlin; info8[2] = info.ndiag; info8[3] = info.ndense;
        info8[4] = info.conflicts; info8[5] = info.vec_rounds; info8[6] = info.fma_per_amp; info8[7] = info.mat_elems;
    }
    return rc;
}

extern "C" __attribute__((visibility("default"))) int emu_apply_gate_pass(void* state, int nbits, int dtype, int nops, const int* ops_k,
                                                                           const int* ops_bits, const double* ops_mats, int n_hi,
                                                                           const int* tile_hi, double* info8) {
    LPassInfo info;
    memset(&info, 0, sizeof(info));
    int rc;
    if (dtype == TCB200_C64) rc = emu_lpass<float>(state, nbits, nops, ops_k, ops_bits, ops_mats, n_hi, tile_hi, &info);
    else rc = emu_lpass<double>(state, nbits, nops, ops_k, ops_bits, ops_mats, n_hi, tile_hi, &info);
    if (info8) {
        info8[0] = info.rounds; info8[1] = info.nlin; info8[2] = info.ndiag; info8[3] = info.ndense;
        info8[4] = info.conflicts; info8[5] = info.vec_rounds; info8[6] = info.fma_per_amp; info8[7] = info.mat_elems;
    }
    return rc;
}
#endif

}  // namespace tcb

using namespace tcb;

static int gate_pass_args(int nbits, int dtype, int nops, const int* ops_k, const int* ops_bits, const double* ops_mats) {
    if (!ops_k || !ops_bits || !ops_mats) return fail(TCB200_ERR_ARG, "NULL argument");
    if (dtype != TCB200_C64 && dtype != TCB200_C128) return fail(TCB200_ERR_ARG, "bad dtype %d", dtype);
    if (nops < 1 || nops > TCB200_MAX_GATE_PASS_OPS) return fail(TCB200_ERR_ARG, "nops=%d out of range", nops);
    if (nbits < 1 || nbits > 40) return fail(TCB200_ERR_ARG, "nbits=%d out of range", nbits);
    return 0;
}

static void info_to_array(const LPassInfo& info, double* info8) {
    if (!info8) return;
    info8[0] = info.rounds;
    info8[1] = info.nlin;
    info8[2] = info.ndiag;
    info8[3] = info.ndense;
    info8[4] = info.conflicts;
    info8[5] = info.vec_rounds;
    info8[6] = info.fma_per_amp;
    info8[7] = info.mat_elems;
}

extern "C" {

int tcb200_apply_gate_pass(void* state, int nbits, int dtype, int nops, const int* ops_k, const int* ops_bits,
                           const double* ops_mats, int n_hi, const int* tile_hi, int64_t batch, double* info8, void* stream) {
    if (!state) return fail(TCB200_ERR_ARG, "state is NULL");
    int rc = gate_pass_args(nbits, dtype, nops, ops_k, ops_bits, ops_mats);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    LPassInfo info;
    memset(&info, 0, sizeof(info));
    if (dtype == TCB200_C64) rc = launch_lpass<float>(state, nbits, nops, ops_k, ops_bits, ops_mats, n_hi, tile_hi, batch, st, &info);
    else rc = launch_lpass<double>(state, nbits, nops, ops_k, ops_bits, ops_mats, n_hi, tile_hi, batch, st, &info);
    if (rc == 0) info_to_array(info, info8);
    return rc;
}

size_t tcb200_gate_pass_batched_workspace_bytes(int dtype, int64_t batch) {
    (void)dtype;
    return (size_t)(batch < 1 ? 1 : batch) * LP_MAT_ELEMS * 16;
}

int tcb200_apply_gate_pass_batched(void* state, int nbits, int dtype, int nops, const int* ops_k, const int* ops_bits,
                                   const double* ops_mats, const int* ops_batched, int n_hi, const int* tile_hi, int64_t batch,
                                   void* workspace, size_t ws_bytes, double* info8, void* stream) {
    if (!state) return fail(TCB200_ERR_ARG, "state is NULL");
    if (!ops_batched) return fail(TCB200_ERR_ARG, "ops_batched is NULL");
    int rc = gate_pass_args(nbits, dtype, nops, ops_k, ops_bits, ops_mats);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    LPassInfo info;
    memset(&info, 0, sizeof(info));
    if (dtype == TCB200_C64)
        rc = launch_lpass<float>(state, nbits, nops, ops_k, ops_bits, ops_mats, n_hi, tile_hi, batch, st, &info, ops_batched, workspace, ws_bytes);
    else
        rc = launch_lpass<double>(state, nbits, nops, ops_k, ops_bits, ops_mats, n_hi, tile_hi, batch, st, &info, ops_batched, workspace, ws_bytes);
    if (rc == 0) info_to_array(info, info8);
    return rc;
}

int tcb200_gate_pass_info(int nbits, int dtype, int nops, const int* ops_k, const int* ops_bits, const double* ops_mats,
                          int n_hi, const int* tile_hi, double* info8) {
    int rc = gate_pass_args(nbits, dtype, nops, ops_k, ops_bits, ops_mats);
    if (rc) return rc;
    LPassInfo info;
    memset(&info, 0, sizeof(info));
    if (dtype == TCB200_C64) rc = launch_lpass<float>(nullptr, nbits, nops, ops_k, ops_bits, ops_mats, n_hi, tile_hi, 1, nullptr, &info);
    else rc = launch_lpass<double>(nullptr, nbits, nops, ops_k, ops_bits, ops_mats, n_hi, tile_hi, 1, nullptr, &info);
    if (rc == 0) info_to_array(info, info8);
    return rc;
}

}  // extern "C"
