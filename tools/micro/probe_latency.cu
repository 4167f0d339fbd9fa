// probe_latency.cu -- how long does one hash-slot probe (three 16-byte loads inside one random
// 64-byte line, after a 16-byte store to the previous line) take on this GPU, and does an L2
// prefetch issued AHEAD lanes earlier shorten it?  One warp per SM, lanes 0..2 active like the
// codec kernels.  Prints cycles per probe.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o probe_latency probe_latency.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ uint4 ld128(const unsigned char *p) {
    uint4 v;
    asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ld128cg(const unsigned char *p) {
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st128cg(unsigned char *p, uint4 v) {
    asm volatile("st.global.cg.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st128(unsigned char *p, uint4 v) {
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ unsigned rng(unsigned &s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }
// mode 0: plain; 1: prefetch.global.L2 AHEAD probes ahead; 2: ld.cg AHEAD probes ahead (value consumed late);
// 3: no write-back store; 4: store issued after the loads; 5: ld.cg probe loads; 6: st.cg write-back; 7: both .cg;
// 8: both .cg and store after loads
template <int AHEAD>
__global__ void k(unsigned char *buf, u64 lines, int iters, int mode, u64 *out, int warps_per_cta) {
    const int lane = threadIdx.x & 31;
    if (lane >= 3) return;
    unsigned s = 12345u + 977u * (blockIdx.x * 64 + (threadIdx.x >> 5)) + 31u * lane;
    u64 addr[AHEAD + 1];
    for (int i = 0; i <= AHEAD; ++i) addr[i] = (u64(rng(s)) * 2654435761ull % lines) * 64;
    unsigned acc = 0, pend = 0;
    unsigned char *prev = nullptr;
    uint4 sl = make_uint4(1, 2, 3, 4);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        unsigned char *b0 = buf + addr[0];
#pragma unroll
        for (int i = 0; i < AHEAD; ++i) addr[i] = addr[i + 1];
        addr[AHEAD] = (u64(rng(s)) * 2654435761ull % lines) * 64;
        if (mode == 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(buf + addr[AHEAD - 1]));
        if (mode == 2) { acc += pend; asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(pend) : "l"(buf + addr[AHEAD - 1])); }
        if (prev && (mode <= 2 || mode == 5)) *reinterpret_cast<uint4 *>(prev) = sl;
        if (prev && (mode == 6 || mode == 7)) st128cg(prev, sl);
        uint4 a0, a1, a2;
        if (mode == 5 || mode == 7 || mode == 8) a0 = ld128cg(b0), a1 = ld128cg(b0 + 16), a2 = ld128cg(b0 + 32);
        else a0 = ld128(b0), a1 = ld128(b0 + 16), a2 = ld128(b0 + 32);
        if (prev && mode == 4) st128(prev, sl);
        if (prev && mode == 13) { asm volatile("" ::: "memory"); st128(prev, make_uint4(a0.x, a1.x, a2.x, sl.y)); }
        if (prev && mode == 8) st128cg(prev, sl);
        if (prev && mode == 9) { *reinterpret_cast<uint4 *>(prev) = sl; *reinterpret_cast<uint4 *>(prev + 16) = sl; }
        if (prev && mode == 10) { *reinterpret_cast<uint4 *>(prev) = sl; *reinterpret_cast<uint4 *>(prev + 16) = sl;
                                  *reinterpret_cast<uint4 *>(prev + 32) = sl; *reinterpret_cast<uint4 *>(prev + 48) = sl; }
        if (prev && mode == 11) *reinterpret_cast<uint4 *>(buf + ((addr[1] * 7 + 64 * 1237) % (lines * 64))) = sl;
        if (prev && mode == 12) *reinterpret_cast<unsigned *>(prev) = sl.y;
        sl.x = a0.x + a1.x + a2.x + it;   // dependent use, like the slot selection
        // ~600 cycles of dependent ALU work between probes, like the four bits of a nibble
        unsigned w = sl.x;
        for (int q = 0; q < 150; ++q) w = w * 1664525u + 1013904223u;
        sl.y = w;
        s ^= (w & 1u);   // next address depends on the work (no run-ahead by the hardware)
        prev = b0;
    }
    const long long t1 = clock64();
    if (lane == 0) out[blockIdx.x * warps_per_cta + (threadIdx.x >> 5)] = u64(t1 - t0) + (acc == 77 ? 1 : 0) + (sl.y == 3 ? 1 : 0);
}
int main() {
    const size_t big = size_t(12) << 30, small = size_t(32) << 20;
    unsigned char *buf;
    cudaMalloc(&buf, big);
    cudaMemset(buf, 0, big);
    u64 *out;
    cudaMallocManaged(&out, 8 * 148 * 8);
    const int iters = 20000;
    for (int wpc : {1}) for (size_t sz : {small, big}) for (int mode : {0, 3, 4, 8, 13}) {
        k<4><<<148, 32 * wpc>>>(buf, sz / 64, iters, mode, out, wpc);
        cudaDeviceSynchronize();
        double sum = 0;
        for (int i = 0; i < 148 * wpc; ++i) sum += double(out[i]);
        printf("warps/SM %d  buffer %5zu MiB  mode %d (%s): %7.1f cycles per probe iteration\n", wpc, sz >> 20, mode,
               mode == 0 ? "plain" : mode == 1 ? "prefetch.L2 4 ahead" : mode == 2 ? "ld.cg 4 ahead" : mode == 3 ? "no write-back" : mode == 4 ? "store after loads" : mode == 5 ? "ld.cg probe" : mode == 6 ? "st.cg" : mode == 7 ? "ld.cg+st.cg" : mode == 8 ? "ld.cg+st.cg, store after loads" : mode == 9 ? "32-byte sector store" : mode == 10 ? "64-byte line store" : mode == 11 ? "16-byte store to an unrelated line" : mode == 12 ? "4-byte store" : "store after the loads have returned",
               sum / (148 * wpc) / iters);
    }
    {   // the ALU work alone
        k<4><<<148, 32>>>(buf, 1, iters, 3, out, 1);
        cudaDeviceSynchronize();
        printf("same line every time (L1/L2 hit): %7.1f\n", double(out[0]) / iters);
    }
    return 0;
}
