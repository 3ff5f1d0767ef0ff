// gather_rate.cu -- what the B200 memory system delivers for SCATTERED small reads, the access pattern of the homology kernels
// (one or two 32-byte sectors from each of four planes per indel), measured the way bench.py measures its kernels: CUDA events,
// L2 flushed by touching a buffer larger than L2 between repetitions, inputs far larger than L2 (4 GiB).
//
//   pattern "random":    every access is one 8-byte load at a uniformly random 32-byte-aligned offset of the buffer
//   pattern "clustered": consecutive threads read offsets that grow by ~135 bytes with jitter (neighbouring indels of a record in a
//                        2-bit plane at one indel per 540 bases), 4 independent streams per thread (the four planes)
// Every thread issues LOADS independent loads before it uses any of them. Output: one JSON line per configuration with sectors/s and
// GB/s counted as 32-byte sectors and as 64-byte DRAM bursts.
//
// Build (build container):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/microbench/gather_rate.bin profiles/microbench/gather_rate.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

template <int LOADS>
__global__ void __launch_bounds__(256) gather_kernel(const uint64_t *__restrict__ buf, uint64_t n_sectors, int clustered, uint64_t seed,
                                                     uint64_t *__restrict__ out)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t v[LOADS];
#pragma unroll
    for (int i = 0; i < LOADS; i++) {
        uint64_t sector;
        if (clustered) {   // stream i of 4 lives in its own quarter of the buffer; thread t sits ~135 bytes after thread t - 1
            const uint64_t quarter = n_sectors / 4;
            const uint64_t byte = t * 135 + (mix(t * 4 + i + seed) & 63);
            sector = (uint64_t)(i & 3) * quarter + ((byte / 32 + (uint64_t)(i >> 2) * 1000003ull) % quarter);
        } else {
            sector = mix(t * LOADS + i + seed) % n_sectors;
        }
        v[i] = __ldg(buf + sector * 4);
    }
    uint64_t acc = 0;
#pragma unroll
    for (int i = 0; i < LOADS; i++) acc ^= v[i];
    if (acc == 0x1234567887654321ull) out[0] = acc;   // keeps the loads alive
}

template <int LOADS>
static void run(const uint64_t *buf, uint64_t n_sectors, int clustered, uint64_t *out, void *flush, size_t flush_bytes, uint64_t n_threads)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float best = 1e30f, sum = 0;
    const int reps = 5;
    for (int r = 0; r < reps + 1; r++) {
        CK(cudaMemsetAsync(flush, r, flush_bytes));
        CK(cudaEventRecord(a));
        gather_kernel<LOADS><<<(unsigned)(n_threads / 256), 256>>>(buf, n_sectors, clustered, 977 * (r + 1), out);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (r == 0) continue;    // warm-up
        best = ms < best ? ms : best; sum += ms;
    }
    const double acc = (double)n_threads * LOADS, ms = sum / reps;
    printf("{\"pattern\": \"%s\", \"loads_in_flight_per_thread\": %d, \"accesses\": %.0f, \"ms_mean\": %.4f, \"ms_best\": %.4f, "
           "\"gsectors_per_s\": %.2f, \"GBps_32B_sectors\": %.1f, \"GBps_64B_bursts\": %.1f}\n",
           clustered ? "clustered" : "random", LOADS, acc, ms, best, acc / ms / 1e6, acc * 32 / ms / 1e6, acc * 64 / ms / 1e6);
}

int main()
{
    const size_t bytes = (size_t)4 << 30, flush_bytes = (size_t)256 << 20;
    uint64_t *buf, *out; void *flush;
    CK(cudaMalloc(&buf, bytes)); CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&flush, flush_bytes));
    CK(cudaMemset(buf, 1, bytes));
    const uint64_t n_sectors = bytes / 32;
    for (int clustered = 0; clustered < 2; clustered++) {
        const uint64_t n_threads = clustered ? (uint64_t)370433 / 256 * 256 : (uint64_t)1 << 22;   // clustered: one thread per C2 indel
        run<1>(buf, n_sectors, clustered, out, flush, flush_bytes, n_threads);
        run<4>(buf, n_sectors, clustered, out, flush, flush_bytes, n_threads);
        run<8>(buf, n_sectors, clustered, out, flush, flush_bytes, n_threads);
        run<16>(buf, n_sectors, clustered, out, flush, flush_bytes, n_threads);
    }
    // many threads, random: the ceiling with the machine full
    run<4>(buf, n_sectors, 0, out, flush, flush_bytes, (uint64_t)1 << 25);
    run<8>(buf, n_sectors, 0, out, flush, flush_bytes, (uint64_t)1 << 25);
    // streaming reference on the same buffer: cudaMemcpy device to device of 1 GiB
    {
        cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
        CK(cudaMemcpy((char *)buf + ((size_t)2 << 30), buf, (size_t)1 << 30, cudaMemcpyDeviceToDevice));
        CK(cudaEventRecord(a));
        CK(cudaMemcpyAsync((char *)buf + ((size_t)2 << 30), buf, (size_t)1 << 30, cudaMemcpyDeviceToDevice));
        CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        printf("{\"pattern\": \"stream (memcpy D2D 1 GiB, read + write)\", \"ms_mean\": %.4f, \"GBps\": %.1f}\n", ms, 2.0 * (1 << 30) / ms / 1e6);
    }
    return 0;
}
