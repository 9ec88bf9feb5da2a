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
