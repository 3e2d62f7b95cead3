// pg2_microbench.cu -- measures the FP64 issue rate of the device the engine runs on: the denominator of
// the fill kernels' "DP-issue" roofline (the fill is FP64/INT-issue bound, and the B200
// FP64 non-tensor peak is not in MEASURED_PEAKS.json).  Two loops of independent operations per thread,
// enough warps to saturate every SM sub-partition: DADD only, and the DADD + DSETP + select mix of one
// first-wins candidate update.  Results are warp-instructions per second over the whole chip.
#include "../../include/pagan2_b200.h"
#include "pg2_device.cuh"

namespace pg2 {
#ifndef PG2_HOST_EMU
template <int CHAINS>
__global__ void __launch_bounds__(256) dadd_kernel(double *out, double inc, int iters) {
    double a[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) a[k] = (double)(threadIdx.x + k);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < CHAINS; ++k) a[k] = __dadd_rn(a[k], inc);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) s += a[k];
    if (s == 12345.678) out[0] = s;  // keep the loop alive
}

// one candidate update: s = src + c; if (s > best) { best = s; ptr = tag; }
template <int CHAINS>
__global__ void __launch_bounds__(256) candidate_kernel(double *out, double inc, int iters) {
    double best[CHAINS], src[CHAINS];
    unsigned ptr[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) { best[k] = -1e300; src[k] = (double)(threadIdx.x * 3 + k); ptr[k] = 0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < CHAINS; ++k) {
            double s = __dadd_rn(src[k], inc);
            if (s > best[k]) { best[k] = s; ptr[k] = (unsigned)it; }
            src[k] = s;
        }
    }
    double s = 0;
    unsigned p = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) { s += best[k]; p ^= ptr[k]; }
    if (s == 12345.678 || p == 0xdeadbeefu) out[0] = s;
}

// Issue-port check: NF independent FADDs and ND independent DADDs per thread and iteration.  If an FP64 instruction
// held the sub-partition's dispatch port for its two pipe cycles, the mix would cost NF + 2*ND cycles per warp and
// iteration; if other work issues in the shadow of the FP64 pipe it costs max(NF + ND, 2*ND).  Measured on B200:
// 16.1 / 16.5 / 25.6 cycles for (0,8) / (16,0) / (16,8), i.e. the second.
template <int NF, int ND>
__global__ void __launch_bounds__(256) mix_kernel(double *out, double dinc, float finc, int iters) {
    double a[ND > 0 ? ND : 1];
    float f[NF > 0 ? NF : 1];
#pragma unroll
    for (int k = 0; k < ND; ++k) a[k] = (double)(threadIdx.x + k);
#pragma unroll
    for (int k = 0; k < NF; ++k) f[k] = (float)(threadIdx.x + 2 * k);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < ND; ++k) a[k] = __dadd_rn(a[k], dinc);
#pragma unroll
        for (int k = 0; k < NF; ++k) f[k] = __fadd_rn(f[k], finc);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < ND; ++k) s += a[k];
#pragma unroll
    for (int k = 0; k < NF; ++k) s += (double)f[k];
    if (s == 12345.678) out[0] = s;
}
#endif
}  // namespace pg2

// dadd_gips: 1e9 DADD warp-instructions/s chip-wide; cand_gips: 1e9 candidate updates (DADD+DSETP+selects)
// warp-instructions-groups/s; sm_clock_mhz: clock the driver reports during the run (informational).
extern "C" int pg2_measure_fp64_issue(int device, double *dadd_gips, double *cand_gips) {
#ifdef PG2_HOST_EMU
    (void)device; *dadd_gips = 0; *cand_gips = 0;
    return PG2_ERR_NO_DEVICE;
#else
    if (cudaSetDevice(device) != cudaSuccess) return PG2_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PG2_ERR_CUDA;
    double *out = nullptr;
    if (cudaMalloc((void **)&out, 64) != cudaSuccess) return PG2_ERR_NOMEM;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int CH = 8, threads = 256, blocks = prop.multiProcessorCount * 8, iters = 1 << 15;
    float ms = 0;
    double best_d = 0, best_c = 0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        pg2::dadd_kernel<CH><<<blocks, threads>>>(out, 1.0000001, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double warp_instr = (double)blocks * (threads / 32) * CH * (double)iters;
        if (rep && warp_instr / (ms * 1e-3) > best_d) best_d = warp_instr / (ms * 1e-3);
        cudaEventRecord(e0);
        pg2::candidate_kernel<CH><<<blocks, threads>>>(out, 1.0000001, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && warp_instr / (ms * 1e-3) > best_c) best_c = warp_instr / (ms * 1e-3);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    if (cudaGetLastError() != cudaSuccess) return PG2_ERR_CUDA;
    *dadd_gips = best_d * 1e-9;
    *cand_gips = best_c * 1e-9;
    return PG2_OK;
#endif
}

// Cycles per warp and iteration on one SM sub-partition (8 resident warps each) for three loops: 8 DADD, 16 FADD,
// and 8 DADD + 16 FADD (see mix_kernel).
extern "C" int pg2_measure_dispatch_mix(int device, double *cycles, double *sm_clock_mhz) {
#ifdef PG2_HOST_EMU
    (void)device; (void)cycles; (void)sm_clock_mhz;
    return PG2_ERR_NO_DEVICE;
#else
    if (!cycles || !sm_clock_mhz) return PG2_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return PG2_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PG2_ERR_CUDA;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
    *sm_clock_mhz = khz * 1e-3;
    double *out = nullptr;
    if (cudaMalloc((void **)&out, 64) != cudaSuccess) return PG2_ERR_NOMEM;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int threads = 256, blocks = prop.multiProcessorCount * 4, iters = 1 << 15;  // 32 warps per SM = 8 per sub-partition
    for (int which = 0; which < 3; ++which) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            float ms = 0;
            cudaEventRecord(e0);
            if (which == 0) pg2::mix_kernel<0, 8><<<blocks, threads>>>(out, 1.0000001, 1.0001f, iters);
            else if (which == 1) pg2::mix_kernel<16, 0><<<blocks, threads>>>(out, 1.0000001, 1.0001f, iters);
            else pg2::mix_kernel<16, 8><<<blocks, threads>>>(out, 1.0000001, 1.0001f, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep && ms < best) best = ms;
        }
        // warps per sub-partition = blocks * 8 / (SMs * 4) = 8; cycles per warp-iteration of ONE warp's share of the port
        const double total_cycles = best * 1e-3 * khz * 1e3;
        cycles[which] = total_cycles / iters / 8.0;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    if (cudaGetLastError() != cudaSuccess) return PG2_ERR_CUDA;
    return PG2_OK;
#endif
}
