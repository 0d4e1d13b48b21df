// State initialisation, |psi|^2 reductions, probability vector and the two-level CDF sampler.
#include <string.h>

#include "common.cuh"

namespace tcb {

// ------------------------------------------------------------------------------------------------
// init / load
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) zero_kernel(Unit16* v, uint64_t nunits_per_vec, int esz) {
    // every vector: all zero except amplitude 0 = 1
    Unit16* vec = v + (uint64_t)blockIdx.y * nunits_per_vec;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; u < nunits_per_vec; u += stride) {
        Unit16 q;
        q.w[0] = q.w[1] = q.w[2] = q.w[3] = 0u;
        if (u == 0) {
            if (esz == 8) {
                q.w[0] = __float_as_uint(1.0f);
            } else {
                const unsigned long long one = (unsigned long long)__double_as_longlong(1.0);
                q.w[0] = (uint32_t)(one & 0xffffffffull);
                q.w[1] = (uint32_t)(one >> 32);
            }
        }
        vec[u] = q;
    }
}

template <typename Real>
__global__ void __launch_bounds__(256) load_c128_kernel(typename CT<Real>::type* dst, const double2* src, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double2 s = src[i];
        typename CT<Real>::type d;
        d.x = (Real)s.x;
        d.y = (Real)s.y;
        dst[i] = d;
    }
}

static unsigned stream_grid(uint64_t work_items, unsigned per_cta) {
    uint64_t want = (work_items + per_cta - 1) / per_cta;
    if (want < 1) want = 1;
    const uint64_t cap = 148ull * 16;  // multiple of the SM count
    return (unsigned)(want < cap ? want : cap);
}

// ------------------------------------------------------------------------------------------------
// deterministic block sums of |psi|^2
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sum over the CTA (blockDim.x multiple of 32, <= 1024); result valid in thread 0
__device__ __forceinline__ double cta_sum(double v, double* sh /*[32]*/) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        const int nw = blockDim.x >> 5;
        r = l < nw ? sh[l] : 0.0;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}

template <typename Real>
__device__ __forceinline__ double unit_prob(const Unit16& q) {
    using C = typename CT<Real>::type;
    const C* c = reinterpret_cast<const C*>(&q);
    double p = 0.0;
#pragma unroll
    for (int a = 0; a < CT<Real>::APU; ++a) p += (double)c[a].x * (double)c[a].x + (double)c[a].y * (double)c[a].y;
    return p;
}

// out[y][b] = sum of |psi|^2 over block b (2^B amplitudes) of vector y
template <typename Real>
__global__ void __launch_bounds__(256) block_sums_kernel(const Unit16* state, int n, int B, double* out) {
    __shared__ double sh[32];
    constexpr int APU = CT<Real>::APU;
    const uint64_t units_per_vec = (1ull << n) / APU;
    const uint32_t units_per_block = (1u << B) / APU;
    const uint64_t nb = 1ull << (n - B);
    const Unit16* vec = state + (uint64_t)blockIdx.y * units_per_vec;
    for (uint64_t b = blockIdx.x; b < nb; b += gridDim.x) {
        const Unit16* blk = vec + b * units_per_block;
        double acc = 0.0;
        for (uint32_t u = threadIdx.x; u < units_per_block; u += blockDim.x) acc += unit_prob<Real>(blk[u]);
        const double s = cta_sum(acc, sh);
        if (threadIdx.x == 0) out[(uint64_t)blockIdx.y * nb + b] = s;
    }
}

// out[y] = sum_i in[y][i], single CTA per y, fixed order
__global__ void __launch_bounds__(1024) final_sum_kernel(const double* in, uint64_t count, double* out) {
    __shared__ double sh[32];
    const double* v = in + (uint64_t)blockIdx.x * count;
    double acc = 0.0;
    for (uint64_t i = threadIdx.x; i < count; i += blockDim.x) acc += v[i];
    const double s = cta_sum(acc, sh);
    if (threadIdx.x == 0) out[blockIdx.x] = s;
}

// out_e = (|in_e|^2, 0)  (mode 0)   or   out_e = (sqrt(max(Re in_e, 0)), 0)  (mode 1); in may equal out
template <typename Real>
__global__ void __launch_bounds__(256) prob_state_kernel(const typename CT<Real>::type* in, typename CT<Real>::type* out,
                                                         uint64_t n, int mode) {
    using C = typename CT<Real>::type;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const C c = in[i];
        C o;
        o.x = mode == 0 ? c.x * c.x + c.y * c.y : sqrt(c.x > (Real)0 ? c.x : (Real)0);
        o.y = 0;
        out[i] = o;
    }
}

// part[c] = sum of |psi_e|^2 over the amplitudes of CTA c's grid-stride share with (e & mask) == value:
// the probability mass of a partial measurement record (perfect sampling, basecircuit.py:359-443)
template <typename Real>
__global__ void __launch_bounds__(256) masked_sums_kernel(const Unit16* state, int n, uint64_t mask, uint64_t value, double* part) {
    __shared__ double sh[32];
    constexpr int APU = CT<Real>::APU;
    using C = typename CT<Real>::type;
    const uint64_t units = (1ull << n) / APU > 0 ? (1ull << n) / APU : 1;
    double acc = 0.0;
    for (uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; u < units; u += (uint64_t)gridDim.x * blockDim.x) {
        const Unit16 q = state[u];
        const C* a = reinterpret_cast<const C*>(&q);
#pragma unroll
        for (int j = 0; j < APU; ++j) {
            const uint64_t e = u * APU + j;
            if ((e & mask) == value && e < (1ull << n)) acc += (double)a[j].x * (double)a[j].x + (double)a[j].y * (double)a[j].y;
        }
    }
    const double s = cta_sum(acc, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
}

template <typename Real>
__global__ void __launch_bounds__(256) prob_kernel(const typename CT<Real>::type* state, Real* out, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const typename CT<Real>::type c = state[i];
        out[i] = c.x * c.x + c.y * c.y;
    }
}

static int reduce_block_bits(int nbits) {
    int B = nbits < 12 ? nbits : 12;
    if (nbits - B > 22) B = nbits - 22;
    return B;
}

// ------------------------------------------------------------------------------------------------
// inclusive scan of the block sums (double), three small kernels
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_PER_THREAD = 4;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_PER_THREAD;  // 1024

// exclusive scan of one value per thread over the CTA; returns exclusive prefix, total in *tot
__device__ __forceinline__ double cta_exclusive_scan(double v, double* sh /*[33]*/, double* tot) {
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    double inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, inc, o);
        if (l >= o) inc += t;
    }
    if (l == 31) sh[w] = inc;
    __syncthreads();
    if (w == 0) {
        const int nw = blockDim.x >> 5;
        double x = l < nw ? sh[l] : 0.0;
        double xi = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, xi, o);
            if (l >= o) xi += t;
        }
        if (l < nw) sh[l] = xi - x;  // exclusive warp offsets
        if (l == 31) sh[32] = xi;    // total (lanes >= nw contribute 0)
    }
    __syncthreads();
    const double r = sh[w] + (inc - v);
    *tot = sh[32];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_chunks_kernel(const double* in, double* out, double* totals, uint64_t count) {
    __shared__ double sh[33];
    const uint64_t i0 = (uint64_t)blockIdx.x * SCAN_CHUNK + (uint64_t)threadIdx.x * SCAN_PER_THREAD;
    double s[SCAN_PER_THREAD];
    double run = 0.0;
#pragma unroll
    for (int j = 0; j < SCAN_PER_THREAD; ++j) {
        const double x = (i0 + j < count) ? in[i0 + j] : 0.0;
        run += x;
        s[j] = run;
    }
    double tot;
    const double ex = cta_exclusive_scan(run, sh, &tot);
#pragma unroll
    for (int j = 0; j < SCAN_PER_THREAD; ++j)
        if (i0 + j < count) out[i0 + j] = ex + s[j];
    if (threadIdx.x == 0) totals[blockIdx.x] = tot;
}

// in-place exclusive scan of up to 4096 totals by a single CTA of 1024 threads
__global__ void __launch_bounds__(1024) scan_totals_kernel(double* totals, int count) {
    __shared__ double sh[33];
    const int i0 = threadIdx.x * 4;
    double s[4];
    double run = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double x = (i0 + j < count) ? totals[i0 + j] : 0.0;
        s[j] = run;  // exclusive within the thread
        run += x;
    }
    double tot;
    const double ex = cta_exclusive_scan(run, sh, &tot);
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (i0 + j < count) totals[i0 + j] = ex + s[j];
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_add_kernel(double* out, const double* offsets, uint64_t count) {
    const double off = offsets[blockIdx.x];
    const uint64_t i0 = (uint64_t)blockIdx.x * SCAN_CHUNK;
    for (int j = threadIdx.x; j < SCAN_CHUNK; j += SCAN_THREADS)
        if (i0 + j < count) out[i0 + j] += off;
}

// ------------------------------------------------------------------------------------------------
// per-shot search: one warp per shot
// ------------------------------------------------------------------------------------------------
template <typename Real>
__global__ void __launch_bounds__(256) sample_search_kernel(const typename CT<Real>::type* state, int n, int B, const double* bcdf,
                                                            const double* uniforms, int64_t shots, int64_t* out,
                                                            double cdf_offset, double cdf_total) {
    using C = typename CT<Real>::type;
    const int lane = threadIdx.x & 31;
    const int64_t shot = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (shot >= shots) return;
    const uint64_t nb = 1ull << (n - B);
    const double local_total = bcdf[nb - 1];
    const double u = uniforms[shot];
    double r;
    if (cdf_total < 0.0) {
        r = local_total * (1.0 - u);
        if (r > local_total) r = local_total;
    } else {
        r = cdf_total * (1.0 - u) - cdf_offset;
        if (!(r > 0.0) || r > local_total) {
            if (lane == 0) out[shot] = -1;
            return;
        }
    }
    // first block b with bcdf[b] >= r
    uint64_t lo = 0, hi = nb - 1;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (bcdf[mid] >= r) hi = mid; else lo = mid + 1;
    }
    const uint64_t b = lo;
    double start = b > 0 ? bcdf[b - 1] : 0.0;
    const C* blk = state + (b << B);
    const uint32_t bs = 1u << B;
    int64_t found = -1;
    // 4 consecutive amplitudes per lane and iteration (128 per warp step): one warp scan per 128
    // amplitudes, then the owning lane resolves its 4 entries
    for (uint32_t i0 = 0; i0 < bs; i0 += 128) {
        const uint32_t i = i0 + 4u * lane;
        double p[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            p[j] = 0.0;
            if (i + j < bs) {
                const C c = blk[i + j];
                p[j] = (double)c.x * (double)c.x + (double)c.y * (double)c.y;
            }
        }
        const double s = (p[0] + p[1]) + (p[2] + p[3]);
        double inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        const double cum = start + inc;
        const unsigned m = __ballot_sync(0xffffffffu, (cum >= r) && (i < bs));
        if (m) {
            const int L = __ffs(m) - 1;
            int64_t mine = -1;
            if (lane == L) {
                double run = cum - s;  // mass before this lane's first entry
                int j = 0;
                for (; j < 3; ++j) {
                    run += p[j];
                    if (run >= r) break;
                }
                uint32_t idx = i + (uint32_t)j;
                if (idx >= bs) idx = bs - 1;
                mine = (int64_t)((b << B) + idx);
            }
            found = __shfl_sync(0xffffffffu, mine, L);
            break;
        }
        start = __shfl_sync(0xffffffffu, cum, 31);
    }
    if (found < 0) found = (int64_t)((b << B) + bs - 1);  // rounding-level miss: clamp to the block end
    if (lane == 0) out[shot] = found;
}

}  // namespace tcb

using namespace tcb;

extern "C" {

int tcb200_init_zero(void* state, int nbits, int dtype, int64_t batch, void* stream) {
    if (!state) return fail(TCB200_ERR_ARG, "state is NULL");
    if (dtype != TCB200_C64 && dtype != TCB200_C128) return fail(TCB200_ERR_ARG, "bad dtype %d", dtype);
    if (nbits < 1 || nbits > 40) return fail(TCB200_ERR_ARG, "nbits=%d out of range", nbits);
    if (batch < 1 || batch > 65535) return fail(TCB200_ERR_ARG, "batch=%lld out of range", (long long)batch);
    const int esz = dtype == TCB200_C64 ? 8 : 16;
    const uint64_t nunits = ((uint64_t)esz << nbits) / 16;
    dim3 grid(stream_grid(nunits, 256 * 4), (unsigned)batch);
    zero_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<Unit16*>(state), nunits, esz);
    TCB_LAUNCH_CHECK("zero_kernel");
    return 0;
}

int tcb200_set_zero(void* state, int nbits, int dtype, int64_t batch, void* stream) {
    if (!state) return fail(TCB200_ERR_ARG, "state is NULL");
    if (dtype != TCB200_C64 && dtype != TCB200_C128) return fail(TCB200_ERR_ARG, "bad dtype %d", dtype);
    if (nbits < 1 || nbits > 40 || batch < 1) return fail(TCB200_ERR_ARG, "bad size");
    const size_t esz = dtype == TCB200_C64 ? 8 : 16;
    TCB_CUDA(cudaMemsetAsync(state, 0, (esz << nbits) * (size_t)batch, static_cast<cudaStream_t>(stream)));
    return 0;
}

int tcb200_copy_rows(void* dst, size_t dst_pitch, const void* src, size_t src_pitch,
                     size_t row_bytes, size_t nrows, void* stream) {
    if (!dst || !src) return fail(TCB200_ERR_ARG, "NULL argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    for (size_t r = 0; r < nrows; ++r)
        TCB_CUDA(cudaMemcpyAsync(static_cast<char*>(dst) + r * dst_pitch, static_cast<const char*>(src) + r * src_pitch,
                                 row_bytes, cudaMemcpyDeviceToDevice, st));
    return 0;
}

int tcb200_load_c128(void* state, int nbits, int dtype, const void* src, void* stream) {
    if (!state || !src) return fail(TCB200_ERR_ARG, "NULL argument");
    if (dtype != TCB200_C64 && dtype != TCB200_C128) return fail(TCB200_ERR_ARG, "bad dtype %d", dtype);
    if (nbits < 1 || nbits > 40) return fail(TCB200_ERR_ARG, "nbits=%d out of range", nbits);
    const uint64_t n = 1ull << nbits;
    const unsigned grid = stream_grid(n, 256 * 4);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == TCB200_C64)
        load_c128_kernel<float><<<grid, 256, 0, st>>>(static_cast<float2*>(state), static_cast<const double2*>(src), n);
    else
        load_c128_kernel<double><<<grid, 256, 0, st>>>(static_cast<double2*>(state), static_cast<const double2*>(src), n);
    TCB_LAUNCH_CHECK("load_c128_kernel");
    return 0;
}

size_t tcb200_reduce_workspace_bytes(int nbits, int64_t batch) {
    const int B = reduce_block_bits(nbits);
    return sizeof(double) * ((size_t)batch << (nbits - B)) + 256;
}

int tcb200_norm2(const void* state, int nbits, int dtype, int64_t batch, double* out_dev,
                 void* workspace, size_t ws_bytes, void* stream) {
    if (!state || !out_dev || !workspace) return fail(TCB200_ERR_ARG, "NULL argument");
    if (dtype != TCB200_C64 && dtype != TCB200_C128) return fail(TCB200_ERR_ARG, "bad dtype %d", dtype);
    if (nbits < 1 || nbits > 40) return fail(TCB200_ERR_ARG, "nbits=%d out of range", nbits);
    if (batch < 1 || batch > 65535) return fail(TCB200_ERR_ARG, "batch=%lld out of range", (long long)batch);
    if (ws_bytes < tcb200_reduce_workspace_bytes(nbits, batch)) return fail(TCB200_ERR_WORKSPACE, "workspace too small");
    const int B = reduce_block_bits(nbits);
    const uint64_t nb = 1ull << (nbits - B);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double* part = static_cast<double*>(workspace);
    dim3 grid((unsigned)(nb < 148ull * 16 ? nb : 148ull * 16), (unsigned)batch);
    if (dtype == TCB200_C64)
        block_sums_kernel<float><<<grid, 256, 0, st>>>(static_cast<const Unit16*>(state), nbits, B, part);
    else
        block_sums_kernel<double><<<grid, 256, 0, st>>>(static_cast<const Unit16*>(state), nbits, B, part);
    TCB_LAUNCH_CHECK("block_sums_kernel");
    final_sum_kernel<<<(unsigned)batch, 1024, 0, st>>>(part, nb, out_dev);
    TCB_LAUNCH_CHECK("final_sum_kernel");
    return 0;
}

size_t tcb200_masked_norm2_workspace_bytes(void) { return sizeof(double) * 148 * 8 + 256; }

int tcb200_masked_norm2(const void* state, int nbits, int dtype, uint64_t mask, uint64_t value, double* out_dev,
                        void* workspace, size_t ws_bytes, void* stream) {
    if (!state || !out_dev || !workspace) return fail(TCB200_ERR_ARG, "NULL argument");
    if (dtype != TCB200_C64 && dtype != TCB200_C128) return fail(TCB200_ERR_ARG, "bad dtype %d", dtype);
    if (nbits < 1 || nbits > 40) return fail(TCB200_ERR_ARG, "nbits=%d out of range", nbits);
    if ((mask >> nbits) || (value & ~mask)) return fail(TCB200_ERR_ARG, "mask / value outside the state or inconsistent");
    if (ws_bytes < tcb200_masked_norm2_workspace_bytes()) return fail(TCB200_ERR_WORKSPACE, "workspace too small");
    const uint64_t units = (1ull << nbits) >> (dtype == TCB200_C64 ? 1 : 0);
    uint64_t want = (units + 255) / 256;
    if (want < 1) want = 1;
    const unsigned grid = (unsigned)(want < 148ull * 8 ? want : 148ull * 8);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double* part = static_cast<double*>(workspace);
    if (dtype == TCB200_C64) {
        masked_sums_kernel<float><<<grid, 256, 0, st>>>(static_cast<const Unit16*>(state), nbits, mask, value, part);
    } else {
        masked_sums_kernel<double><<<grid, 256, 0, st>>>(static_cast<const Unit16*>(state), nbits, mask, value, part);
    }
    TCB_LAUNCH_CHECK("masked_sums_kernel");
    final_sum_kernel<<<1, 1024, 0, st>>>(part, grid, out_dev);
    TCB_LAUNCH_CHECK("final_sum_kernel");
    return 0;
}

int tcb200_probability_state(const void* in, void* out, int nbits, int dtype, int mode, void* stream) {
    if (!in || !out) return fail(TCB200_ERR_ARG, "NULL argument");
    if (dtype != TCB200_C64 && dtype != TCB200_C128) return fail(TCB200_ERR_ARG, "bad dtype %d", dtype);
    if (nbits < 1 || nbits > 40) return fail(TCB200_ERR_ARG, "nbits=%d out of range", nbits);
    if (mode != 0 && mode != 1) return fail(TCB200_ERR_ARG, "mode=%d", mode);
    const uint64_t n = 1ull << nbits;
    const unsigned grid = stream_grid(n, 256 * 4);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == TCB200_C64)
        prob_state_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float2*>(in), static_cast<float2*>(out), n, mode);
    else
        prob_state_kernel<double><<<grid, 256, 0, st>>>(static_cast<const double2*>(in), static_cast<double2*>(out), n, mode);
    TCB_LAUNCH_CHECK("prob_state_kernel");
    return 0;
}

int tcb200_probability(const void* state, int nbits, int dtype, void* prob_dev, int64_t batch, void* stream) {
    if (!state || !prob_dev) return fail(TCB200_ERR_ARG, "NULL argument");
    if (dtype != TCB200_C64 && dtype != TCB200_C128) return fail(TCB200_ERR_ARG, "bad dtype %d", dtype);
    if (nbits < 1 || nbits > 40 || batch < 1) return fail(TCB200_ERR_ARG, "bad size");
    const uint64_t n = (uint64_t)batch << nbits;
    const unsigned grid = stream_grid(n, 256 * 4);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == TCB200_C64)
        prob_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float2*>(state), static_cast<float*>(prob_dev), n);
    else
        prob_kernel<double><<<grid, 256, 0, st>>>(static_cast<const double2*>(state), static_cast<double*>(prob_dev), n);
    TCB_LAUNCH_CHECK("prob_kernel");
    return 0;
}

size_t tcb200_sample_workspace_bytes(int nbits) {
    const int B = reduce_block_bits(nbits);
    const size_t nb = (size_t)1 << (nbits - B);
    const size_t nchunks = (nb + SCAN_CHUNK - 1) / SCAN_CHUNK;
    return sizeof(double) * (2 * nb + nchunks) + 512;
}

int tcb200_sample(const void* state, int nbits, int dtype, const double* uniforms_dev,
                  int64_t shots, int64_t* out_idx_dev, double* total_dev, double cdf_offset,
                  double cdf_total, void* workspace, size_t ws_bytes, void* stream) {
    if (!state || !workspace) return fail(TCB200_ERR_ARG, "NULL argument");
    if (shots > 0 && (!uniforms_dev || !out_idx_dev)) return fail(TCB200_ERR_ARG, "NULL argument");
    if (dtype != TCB200_C64 && dtype != TCB200_C128) return fail(TCB200_ERR_ARG, "bad dtype %d", dtype);
    if (nbits < 1 || nbits > 40) return fail(TCB200_ERR_ARG, "nbits=%d out of range", nbits);
    if (shots < 0) return fail(TCB200_ERR_ARG, "shots < 0");
    if (ws_bytes < tcb200_sample_workspace_bytes(nbits)) return fail(TCB200_ERR_WORKSPACE, "workspace too small");
    const int B = reduce_block_bits(nbits);
    const uint64_t nb = 1ull << (nbits - B);
    const uint64_t nchunks = (nb + SCAN_CHUNK - 1) / SCAN_CHUNK;
    if (nchunks > 4096) return fail(TCB200_ERR_UNSUPPORTED, "state too large for the sampler scan");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double* bsum = static_cast<double*>(workspace);
    double* bcdf = bsum + nb;
    double* totals = bcdf + nb;
    dim3 grid((unsigned)(nb < 148ull * 16 ? nb : 148ull * 16), 1);
    if (dtype == TCB200_C64)
        block_sums_kernel<float><<<grid, 256, 0, st>>>(static_cast<const Unit16*>(state), nbits, B, bsum);
    else
        block_sums_kernel<double><<<grid, 256, 0, st>>>(static_cast<const Unit16*>(state), nbits, B, bsum);
    TCB_LAUNCH_CHECK("block_sums_kernel");
    scan_chunks_kernel<<<(unsigned)nchunks, SCAN_THREADS, 0, st>>>(bsum, bcdf, totals, nb);
    TCB_LAUNCH_CHECK("scan_chunks_kernel");
    if (nchunks > 1) {
        scan_totals_kernel<<<1, 1024, 0, st>>>(totals, (int)nchunks);
        TCB_LAUNCH_CHECK("scan_totals_kernel");
        scan_add_kernel<<<(unsigned)nchunks, SCAN_THREADS, 0, st>>>(bcdf, totals, nb);
        TCB_LAUNCH_CHECK("scan_add_kernel");
    }
    if (total_dev) TCB_CUDA(cudaMemcpyAsync(total_dev, bcdf + (nb - 1), sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (shots > 0) {
        const int warps_per_cta = 8;
        const uint64_t ctas = ((uint64_t)shots + warps_per_cta - 1) / warps_per_cta;
        if (ctas > 0x7fffffffull) return fail(TCB200_ERR_UNSUPPORTED, "too many shots for one launch");
        if (dtype == TCB200_C64)
            sample_search_kernel<float><<<(unsigned)ctas, warps_per_cta * 32, 0, st>>>(static_cast<const float2*>(state), nbits, B, bcdf, uniforms_dev, shots, out_idx_dev, cdf_offset, cdf_total);
        else
            sample_search_kernel<double><<<(unsigned)ctas, warps_per_cta * 32, 0, st>>>(static_cast<const double2*>(state), nbits, B, bcdf, uniforms_dev, shots, out_idx_dev, cdf_offset, cdf_total);
        TCB_LAUNCH_CHECK("sample_search_kernel");
    }
    return 0;
}

}  // extern "C"
