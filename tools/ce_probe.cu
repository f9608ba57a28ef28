// ce_probe.cu -- how copies and kernels of different streams interact on this box while one stream keeps the host->device
// link saturated (the situation of the host pipeline: ASCII chunk copies of sub-batch k+1 back to back while sub-batch k
// is computed and its results are copied back).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/ce_probe tools/ce_probe.cu && build/ce_probe
// Prints, for each scenario, how long a victim operation in stream B takes from its submission while stream A is flooded.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <atomic>
#include <vector>
#include <cuda_runtime.h>

#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

__global__ void spin_kernel(unsigned long long cycles, int *out)
{
    const unsigned long long t0 = clock64();
    while (clock64() - t0 < cycles) {}
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = 1;
}

int main()
{
    const size_t chunk = 8u << 20, nchunk = 64, big = 256u << 20;
    char *h_src, *h_dst, *d_a, *d_b;
    int *d_flag;
    CU(cudaMallocHost(&h_src, chunk * nchunk));
    CU(cudaMallocHost(&h_dst, big));
    CU(cudaMalloc(&d_a, chunk * nchunk));
    CU(cudaMalloc(&d_b, big));
    CU(cudaMalloc(&d_flag, 4));
    cudaStream_t A, B, C;
    CU(cudaStreamCreateWithFlags(&A, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&B, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&C, cudaStreamNonBlocking));
    cudaEvent_t ev[16], done;
    for (auto &e : ev) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
    int dev_clock_khz = 0;
    CU(cudaDeviceGetAttribute(&dev_clock_khz, cudaDevAttrClockRate, 0));
    const unsigned long long ms5 = (unsigned long long)dev_clock_khz * 5ull;   // ~5 ms of spinning

    std::atomic<bool> stop{false};
    std::atomic<size_t> moved{0};
    auto flood = [&](int depth) {   // stream A: 8 MB host->device copies back to back, `depth` in flight, polled like the feeder does
        size_t issued = 0;
        while (!stop.load()) {
            if (issued >= (size_t)depth) {
                while (cudaEventQuery(ev[issued % depth]) == cudaErrorNotReady) {
                    if (stop.load()) return;
                }
            }
            const size_t k = issued % nchunk;
            cudaMemcpyAsync(d_a + k * chunk, h_src + k * chunk, chunk, cudaMemcpyHostToDevice, A);
            cudaEventRecord(ev[issued % depth], A);
            issued++;
            moved += chunk;
        }
    };

    auto scenario = [&](const char *name, int depth, auto victim) {
        stop = false;
        moved = 0;
        std::thread th;
        const double t_flood = now_ms();
        if (depth > 0) th = std::thread(flood, depth);
        std::this_thread::sleep_for(std::chrono::milliseconds(20));
        const double t0 = now_ms();
        victim();
        // poll like the pipeline does
        double t_ready = -1;
        while (true) {
            cudaError_t q = cudaEventQuery(done);
            if (q == cudaSuccess) { t_ready = now_ms(); break; }
            if (now_ms() - t0 > 400) break;
        }
        const double t_poll_end = now_ms();
        stop = true;
        if (th.joinable()) th.join();
        CU(cudaDeviceSynchronize());
        const double flood_gbs = moved.load() / ((t_poll_end - t_flood) * 1e-3) / 1e9;
        if (t_ready > 0) printf("%-64s ready after %7.2f ms   (flood %.1f GB/s)\n", name, t_ready - t0, flood_gbs);
        else printf("%-64s NOT ready after 400 ms of flooding (flood %.1f GB/s)\n", name, flood_gbs);
    };

    // ---- how fast does ONE stream move back-to-back copies of a given size, issued by 1 or 8 host threads?
    for (int nthreads : {1, 8}) {
        for (size_t sz : {(size_t)1 << 20, (size_t)2 << 20, (size_t)4 << 20, (size_t)8 << 20, (size_t)32 << 20}) {
            const size_t total = (size_t)2 << 30;
            const size_t ncopies = total / sz;
            CU(cudaDeviceSynchronize());
            const double t0 = now_ms();
            std::atomic<size_t> next{0};
            std::vector<std::thread> ths;
            for (int t = 0; t < nthreads; t++)
                ths.emplace_back([&]() {
                    while (true) {
                        const size_t k = next.fetch_add(1);
                        if (k >= ncopies) break;
                        const size_t off = (k * sz) % (chunk * nchunk - sz + 1);
                        cudaMemcpyAsync(d_a + off, h_src + off, sz, cudaMemcpyHostToDevice, A);
                    }
                });
            for (auto &t : ths) t.join();
            const double t_issued = now_ms();
            CU(cudaStreamSynchronize(A));
            const double t1 = now_ms();
            printf("one stream, %d issuing thread(s), %4zu MB copies: %6.1f GB/s  (issue %.1f ms, total %.1f ms, %.1f us per copy)\n", nthreads, sz >> 20,
                   total / ((t1 - t0) * 1e-3) / 1e9, t_issued - t0, t1 - t0, (t1 - t0) * 1e3 / ncopies);
        }
    }
    for (int depth : {0, 8}) {
        printf("---- stream A flood depth %d\n", depth);
        scenario("kernel(5ms) in B", depth, [&] { spin_kernel<<<148, 128, 0, B>>>(ms5, d_flag); cudaEventRecord(done, B); });
        scenario("H2D 8MB in B, then kernel(5ms) in B", depth, [&] {
            cudaMemcpyAsync(d_b, h_src, chunk, cudaMemcpyHostToDevice, B);
            spin_kernel<<<148, 128, 0, B>>>(ms5, d_flag);
            cudaEventRecord(done, B);
        });
        scenario("kernel(5ms) in B, then D2H 1MB in B", depth, [&] {
            spin_kernel<<<148, 128, 0, B>>>(ms5, d_flag);
            cudaMemcpyAsync(h_dst, d_b, 1u << 20, cudaMemcpyDeviceToHost, B);
            cudaEventRecord(done, B);
        });
        scenario("D2H 256MB in B", depth, [&] { cudaMemcpyAsync(h_dst, d_b, big, cudaMemcpyDeviceToHost, B); cudaEventRecord(done, B); });
        scenario("kernel(5ms) in B, then D2H 1MB in C after an event", depth, [&] {
            spin_kernel<<<148, 128, 0, B>>>(ms5, d_flag);
            cudaEventRecord(ev[15], B);
            cudaStreamWaitEvent(C, ev[15], 0);
            cudaMemcpyAsync(h_dst, d_b, 1u << 20, cudaMemcpyDeviceToHost, C);
            cudaEventRecord(done, C);
        });
        scenario("H2D 8MB in A (same stream as the flood), then kernel in B after an event", depth, [&] {
            cudaMemcpyAsync(d_b, h_src, chunk, cudaMemcpyHostToDevice, A);
            cudaEventRecord(ev[14], A);
            cudaStreamWaitEvent(B, ev[14], 0);
            spin_kernel<<<148, 128, 0, B>>>(ms5, d_flag);
            cudaEventRecord(done, B);
        });
    }
    return 0;
}
