// Host <-> device copy ceiling of the box, per device alone and with every device copying at once.
// Build: nvcc -O2 -std=c++17 -o tools/ubench_pcie tools/ubench_pcie.cu -lpthread
// Run  : tools/ubench_pcie                 one process, one thread per device, sweep (devices x allocation x direction)
//        tools/ubench_pcie --procs         one PROCESS per device (forked before any CUDA call), cudaMallocHost buffers
//        options: --mb N (MB per device per repetition, default 512)  --chunk-mb N (one cudaMemcpyAsync per chunk, default 64)
//                 --reps N (default 4)  --affinity (pin worker i to its own slice of the CPUs)
// Why: the end-to-end path of tmpc_solve(TMPC_MEM_HOST) returns 648 B per quadrotor solve; with eight ranks on one host
// the per-rank D2H rate collapses (round 1: 55 GB/s alone, 11 GB/s each with eight).  This tool measures that ceiling
// without any of the library in the way, and whether the allocation path (cudaMallocHost vs cudaHostRegister on
// transparent-huge-page memory), threads vs processes, or CPU affinity move it.
#include <cuda_runtime.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

struct Shared {                      // lives in MAP_SHARED memory so that forked workers can meet
    std::atomic<int> arrive[64];
    double seconds[16];
};

static void barrier(Shared *sh, int slot, int n)
{
    sh->arrive[slot].fetch_add(1);
    while (sh->arrive[slot].load() < n) { }
}

enum Alloc { MALLOCHOST = 0, REGISTER_THP = 1, REGISTER_4K = 2 };
enum Dir { D2H = 0, H2D = 1, BOTH = 2 };
static const char *alloc_name[] = {"cudaMallocHost", "cudaHostRegister(THP)", "cudaHostRegister(4K)"};
static const char *dir_name[] = {"D2H", "H2D", "D2H+H2D"};

static void *host_buf(size_t bytes, int kind)
{
    void *p = nullptr;
    if (kind == MALLOCHOST) { CK(cudaHostAlloc(&p, bytes, cudaHostAllocPortable)); return p; }
    p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) { perror("mmap"); exit(1); }
    madvise(p, bytes, kind == REGISTER_THP ? MADV_HUGEPAGE : MADV_NOHUGEPAGE);
    memset(p, 1, bytes);
    CK(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return p;
}
static void host_free(void *p, size_t bytes, int kind)
{
    if (kind == MALLOCHOST) { CK(cudaFreeHost(p)); return; }
    CK(cudaHostUnregister(p));
    munmap(p, bytes);
}

// one worker = one device; returns seconds for reps x bytes in each active direction
static double worker(int dev, int slot, int nworkers, Shared *sh, size_t bytes, size_t chunk, int reps, int kind, int dir, bool affinity, int wi)
{
    if (affinity) {
        const int ncpu = (int)sysconf(_SC_NPROCESSORS_ONLN), per = ncpu / nworkers > 0 ? ncpu / nworkers : 1;
        cpu_set_t set;
        CPU_ZERO(&set);
        for (int c = wi * per; c < (wi + 1) * per && c < ncpu; ++c) CPU_SET(c, &set);
        sched_setaffinity(0, sizeof set, &set);
    }
    CK(cudaSetDevice(dev));
    void *d_a = nullptr, *d_b = nullptr, *h_a = nullptr, *h_b = nullptr;
    CK(cudaMalloc(&d_a, bytes));
    CK(cudaMemset(d_a, 0, bytes));
    h_a = host_buf(bytes, kind);
    if (dir == BOTH) { CK(cudaMalloc(&d_b, bytes)); h_b = host_buf(bytes, kind); }
    cudaStream_t s0, s1;
    CK(cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
    auto pass = [&]() {
        for (size_t o = 0; o < bytes; o += chunk) {
            const size_t n = bytes - o < chunk ? bytes - o : chunk;
            if (dir == D2H || dir == BOTH) CK(cudaMemcpyAsync((char *)h_a + o, (char *)d_a + o, n, cudaMemcpyDeviceToHost, s0));
            if (dir == H2D) CK(cudaMemcpyAsync((char *)d_a + o, (char *)h_a + o, n, cudaMemcpyHostToDevice, s0));
            if (dir == BOTH) CK(cudaMemcpyAsync((char *)d_b + o, (char *)h_b + o, n, cudaMemcpyHostToDevice, s1));
        }
    };
    pass();
    CK(cudaStreamSynchronize(s0)); CK(cudaStreamSynchronize(s1));
    barrier(sh, slot, nworkers);
    const auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; ++r) pass();
    CK(cudaStreamSynchronize(s0)); CK(cudaStreamSynchronize(s1));
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    barrier(sh, slot + 1, nworkers);
    host_free(h_a, bytes, kind);
    if (h_b) host_free(h_b, bytes, kind);
    CK(cudaFree(d_a));
    if (d_b) CK(cudaFree(d_b));
    CK(cudaStreamDestroy(s0)); CK(cudaStreamDestroy(s1));
    return sec;
}

static void report(const char *how, int n, int kind, int dir, size_t bytes, int reps, const double *sec, bool affinity, size_t chunk)
{
    double worst = 0, sum = 0;
    const double per_dir = (double)bytes * reps / 1e9;
    for (int i = 0; i < n; ++i) { worst = sec[i] > worst ? sec[i] : worst; sum += per_dir / sec[i]; }
    const double mult = dir == BOTH ? 2.0 : 1.0;
    printf("%-9s devices %d  %-22s %-8s chunk %4zu MB%s : per device %6.1f GB/s per direction (mean), aggregate %7.1f GB/s%s\n", how, n,
           alloc_name[kind], dir_name[dir], chunk >> 20, affinity ? " affinity" : "", sum / n, mult * per_dir * n / worst,
           dir == BOTH ? " (both directions summed)" : "");
    fflush(stdout);
}

int main(int argc, char **argv)
{
    size_t mb = 512, chunk_mb = 64;
    int reps = 4;
    bool procs = false, affinity = false;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--procs")) procs = true;
        else if (!strcmp(argv[i], "--affinity")) affinity = true;
        else if (!strcmp(argv[i], "--mb") && i + 1 < argc) mb = strtoull(argv[++i], nullptr, 10);
        else if (!strcmp(argv[i], "--chunk-mb") && i + 1 < argc) chunk_mb = strtoull(argv[++i], nullptr, 10);
        else if (!strcmp(argv[i], "--reps") && i + 1 < argc) reps = atoi(argv[++i]);
    }
    const size_t bytes = mb << 20, chunk = chunk_mb << 20;
    Shared *sh = (Shared *)mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    memset((void *)sh, 0, sizeof(Shared));
    int slot = 0;

    if (procs) {
        // device count without creating a CUDA context in the parent: CUDA_VISIBLE_DEVICES or /proc listing is unreliable,
        // so a short-lived child asks the runtime
        int pfd[2];
        if (pipe(pfd) != 0) return 1;
        if (fork() == 0) { int n = 0; cudaGetDeviceCount(&n); if (write(pfd[1], &n, sizeof n) < 0) _exit(1); _exit(0); }
        int ndev = 0;
        if (read(pfd[0], &ndev, sizeof ndev) != (ssize_t)sizeof ndev) return 1;
        wait(nullptr);
        printf("# %d visible devices, %ld CPUs, one PROCESS per device, %zu MB per device per repetition, %d repetitions\n", ndev,
               sysconf(_SC_NPROCESSORS_ONLN), mb, reps);
        for (int dir = D2H; dir <= BOTH; ++dir)
            for (int n = 1; n <= ndev; n *= 2) {
                for (int i = 0; i < n; ++i)
                    if (fork() == 0) {
                        sh->seconds[i] = worker(i, slot, n, sh, bytes, chunk, reps, MALLOCHOST, dir, affinity, i);
                        _exit(0);
                    }
                for (int i = 0; i < n; ++i) wait(nullptr);
                report("processes", n, MALLOCHOST, dir, bytes, reps, sh->seconds, affinity, chunk);
                slot += 2;
            }
        return 0;
    }

    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    printf("# %d visible devices, %ld CPUs, one THREAD per device, %zu MB per device per repetition, %d repetitions\n", ndev,
           sysconf(_SC_NPROCESSORS_ONLN), mb, reps);
    for (int i = 0; i < ndev; ++i) { CK(cudaSetDevice(i)); CK(cudaFree(0)); }
    auto run = [&](int n, int kind, int dir, size_t ch) {
        std::vector<std::thread> th;
        double sec[16] = {0};
        for (int i = 0; i < n; ++i)
            th.emplace_back([&, i] { sec[i] = worker(i, slot, n, sh, bytes, ch, reps, kind, dir, affinity, i); });
        for (auto &t : th) t.join();
        report("threads", n, kind, dir, bytes, reps, sec, affinity, ch);
        slot += 2;
    };
    for (int kind = MALLOCHOST; kind <= REGISTER_4K; ++kind)
        for (int dir = D2H; dir <= BOTH; ++dir)
            for (int n = 1; n <= ndev; n *= 2) {
                if (slot + 2 >= 64) break;
                if (kind == REGISTER_4K && dir != D2H) continue;
                run(n, kind, dir, chunk);
            }
    // chunk size under contention (the library copies 65,536-instance chunks: 31 MB of x, 9 MB of u, 256 KB of iter ...)
    for (size_t ch : {(size_t)256 << 10, (size_t)4 << 20})
        if (slot + 2 < 64) run(ndev, MALLOCHOST, D2H, ch);
    return 0;
}
