// Host-buffer entry point: the call a reference-side binding would make for the whole path.
#include <string.h>

#include "common.cuh"

using namespace tcb;

extern "C" {

int tcb200_run_circuit_host(void* state, int nbits, int dtype, int init_zero, int npasses,
                            const int* pass_k, const int* pass_bits, const double* pass_mats,
                            int64_t shots, const double* uniforms_host, int64_t* out_idx_host,
                            void* workspace, size_t ws_bytes, void* stream) {
    if (!state) return fail(TCB200_ERR_ARG, "state is NULL");
    if (npasses < 0 || (npasses > 0 && (!pass_k || !pass_bits || !pass_mats))) return fail(TCB200_ERR_ARG, "bad pass list");
    if (shots < 0 || (shots > 0 && (!uniforms_host || !out_idx_host))) return fail(TCB200_ERR_ARG, "bad shot buffers");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = 0;
    if (init_zero) {
        rc = tcb200_init_zero(state, nbits, dtype, 1, stream);
        if (rc) return rc;
    }
    const int* b = pass_bits;
    const double* m = pass_mats;
    for (int i = 0; i < npasses; ++i) {
        const int k = pass_k[i];
        if (k < 1 || k > TCB200_MAX_K) return fail(TCB200_ERR_UNSUPPORTED, "pass %d: k=%d", i, k);
        rc = tcb200_apply_dense(state, nbits, dtype, k, b, m, 1, stream);
        if (rc) return rc;
        b += k;
        m += 2ll << (2 * k);
    }
    if (shots > 0) {
        const size_t need = tcb200_sample_workspace_bytes(nbits) + (size_t)shots * 16 + 64;
        if (!workspace || ws_bytes < need) return fail(TCB200_ERR_WORKSPACE, "workspace too small: need %zu bytes", need);
        unsigned char* w = static_cast<unsigned char*>(workspace);
        double* u_dev = reinterpret_cast<double*>(w);
        int64_t* idx_dev = reinterpret_cast<int64_t*>(w + (size_t)shots * 8);
        size_t off = (size_t)shots * 16;
        off = (off + 63) & ~(size_t)63;
        TCB_CUDA(cudaMemcpyAsync(u_dev, uniforms_host, (size_t)shots * 8, cudaMemcpyHostToDevice, st));
        rc = tcb200_sample(state, nbits, dtype, u_dev, shots, idx_dev, nullptr, 0.0, -1.0, w + off, ws_bytes - off, stream);
        if (rc) return rc;
        TCB_CUDA(cudaMemcpyAsync(out_idx_host, idx_dev, (size_t)shots * 8, cudaMemcpyDeviceToHost, st));
    }
    TCB_CUDA(cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"
