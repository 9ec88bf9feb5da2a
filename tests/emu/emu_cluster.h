// tests/emu/emu_cluster.h -- thread-block CLUSTER on the host emulator: one forked process per CTA (so that `static`
// shared variables stay per-CTA), "global memory" and every CTA's dynamic shared memory in one MAP_SHARED arena mapped
// before the fork (same addresses everywhere, hence cluster.map_shared_rank is pointer arithmetic), cluster.sync() on a
// process-shared pthread barrier. Test infrastructure only.
#pragma once
#include "cuda_emu.h"
#include <pthread.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

namespace emu {
struct ClusterShared {
    pthread_barrier_t bar;
    int nctas;
    size_t smemStride;          // bytes of dynamic shared memory reserved per CTA
    char* smemBase;             // nctas x smemStride
};
inline ClusterShared* cl = nullptr;
inline int cl_rank = 0;

struct Arena {
    char* base = nullptr; size_t size = 0, used = 0;
    explicit Arena(size_t bytes) : size(bytes)
    {
        base = (char*)mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
        if (base == MAP_FAILED) { perror("mmap"); exit(2); }
    }
    template <class T> T* alloc(size_t count)
    {
        used = (used + 63) & ~(size_t)63;
        T* p = (T*)(base + used); used += count * sizeof(T);
        if (used > size) { fprintf(stderr, "emu arena exhausted\n"); exit(2); }
        return p;
    }
};

// body(double* dynSmem) runs as every thread of every CTA. Returns 0 when all CTAs exited normally.
template <class F>
int launch_cluster(Arena& arena, int nctas, int nthreads, size_t smemBytes, F&& body)
{
    ClusterShared* sh = arena.alloc<ClusterShared>(1);
    sh->nctas = nctas; sh->smemStride = (smemBytes + 255) & ~(size_t)255;
    sh->smemBase = arena.alloc<char>(sh->smemStride * nctas);
    pthread_barrierattr_t at; pthread_barrierattr_init(&at); pthread_barrierattr_setpshared(&at, PTHREAD_PROCESS_SHARED);
    pthread_barrier_init(&sh->bar, &at, nctas);
    std::vector<pid_t> kids;
    for (int c = 0; c < nctas; c++) {
        pid_t pid = fork();
        if (pid == 0) {
            cl = sh; cl_rank = c;
            gridDim.x = nctas; gridDim.y = gridDim.z = 1;
            double* dyn = (double*)(sh->smemBase + (size_t)c * sh->smemStride);
            launch_cta(nthreads, (unsigned)c, [&] { body(dyn); });
            _exit(0);
        }
        kids.push_back(pid);
    }
    int bad = 0;
    for (pid_t k : kids) { int st = 0; waitpid(k, &st, 0); if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) bad++; }
    return bad;
}
// ---- named cluster bodies ---------------------------------------------------------------------------------------------------
// A harness writes the code one CTA runs as   EMU_CLUSTER_BODY(name) { ... uses ctx, dyn ... }   and launches it with
// EMU_LAUNCH_CLUSTER(arena, nctas, nthreads, smemBytes, name, ctxPointer). Two ways to run it:
//  * default: one forked process per CTA (launch_cluster above);
//  * -DEMU_CLUSTER_THREADS: every CTA in the SAME process, so that ThreadSanitizer also sees the accesses between CTAs (distributed
//    shared memory, exchanges through global memory). The static __shared__ variables still have to be per CTA, so the harness is
//    compiled a second time as a shared library (-DEMU_AS_LIB -shared -fPIC -fvisibility=hidden), the executable copies that file
//    once per CTA and dlopen()s every copy (own statics each); the executable itself contains no kernel body.
typedef void (*emu_body_fn)(void* ctx, double* dyn);
#define EMU_CLUSTER_BODY(name) extern "C" __attribute__((visibility("default"))) void name(void* ctx, double* dyn)

#if defined(EMU_AS_LIB)
}  // namespace emu
extern "C" __attribute__((visibility("default"))) void emu_run_cta(emu::ClusterShared* sh, int rank, int nthreads, emu::emu_body_fn body, void* ctx)
{
    emu::cl = sh; emu::cl_rank = rank;
    gridDim.x = sh->nctas; gridDim.y = gridDim.z = 1;
    double* dyn = (double*)(sh->smemBase + (size_t)rank * sh->smemStride);
    emu::launch_cta(nthreads, (unsigned)rank, [&] { body(ctx, dyn); });
}
namespace emu {
#endif

#if defined(EMU_CLUSTER_THREADS) && !defined(EMU_AS_LIB)
}  // namespace emu
#include <dlfcn.h>
#include <fstream>
#include <string>
namespace emu {
inline std::vector<void*>& cta_libs()
{
    static std::vector<void*> libs;
    return libs;
}
inline void* cta_lib(int rank)
{
    auto& libs = cta_libs();
    while ((int)libs.size() <= rank) {
        const char* src = getenv("EMU_BODY_LIB");
        if (!src) { fprintf(stderr, "EMU_BODY_LIB is not set\n"); exit(2); }
        const std::string dst = std::string(src) + ".cta" + std::to_string(libs.size()) + "." + std::to_string((int)getpid()) + ".so";
        { std::ifstream in(src, std::ios::binary); std::ofstream out(dst, std::ios::binary); out << in.rdbuf(); }
        void* h = dlopen(dst.c_str(), RTLD_NOW | RTLD_LOCAL);
        unlink(dst.c_str());
        if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); exit(2); }
        libs.push_back(h);
    }
    return libs[rank];
}
inline int launch_cluster_named(Arena& arena, int nctas, int nthreads, size_t smemBytes, const char* bodyName, void* ctx)
{
    ClusterShared* sh = arena.alloc<ClusterShared>(1);
    sh->nctas = nctas; sh->smemStride = (smemBytes + 255) & ~(size_t)255;
    sh->smemBase = arena.alloc<char>(sh->smemStride * nctas);
    pthread_barrier_init(&sh->bar, nullptr, nctas);
    typedef void (*run_fn)(ClusterShared*, int, int, emu_body_fn, void*);
    std::vector<std::thread> th;
    for (int c = 0; c < nctas; c++) {
        void* h = cta_lib(c);
        run_fn run = (run_fn)dlsym(h, "emu_run_cta");
        emu_body_fn body = (emu_body_fn)dlsym(h, bodyName);
        if (!run || !body) { fprintf(stderr, "emu: %s not found in the body library\n", bodyName); exit(2); }
        th.emplace_back([=] { run(sh, c, nthreads, body, ctx); });
    }
    for (auto& t : th) t.join();
    pthread_barrier_destroy(&sh->bar);
    return 0;
}
#define EMU_LAUNCH_CLUSTER(arena, nctas, nthreads, smem, name, ctx) emu::launch_cluster_named(arena, nctas, nthreads, smem, #name, ctx)
#else
#define EMU_LAUNCH_CLUSTER(arena, nctas, nthreads, smem, name, ctx) emu::launch_cluster(arena, nctas, nthreads, smem, [&](double* dyn_) { name(ctx, dyn_); })
#endif
}  // namespace emu

namespace cooperative_groups {
struct cluster_group {
    unsigned block_rank() const { return (unsigned)emu::cl_rank; }
    unsigned num_blocks() const { return (unsigned)emu::cl->nctas; }
    void sync() const
    {
        __syncthreads();
        if (threadIdx.x == 0) pthread_barrier_wait(&emu::cl->bar);
        __syncthreads();
    }
    template <class T> T* map_shared_rank(T* p, unsigned rank) const
    {
        char* mine = emu::cl->smemBase + (size_t)emu::cl_rank * emu::cl->smemStride;
        const size_t off = (char*)p - mine;
        if (off >= emu::cl->smemStride) { fprintf(stderr, "emu: map_shared_rank on a pointer outside dynamic shared memory\n"); abort(); }
        return (T*)(emu::cl->smemBase + (size_t)rank * emu::cl->smemStride + off);
    }
};
inline cluster_group this_cluster() { return cluster_group(); }
}  // namespace cooperative_groups
