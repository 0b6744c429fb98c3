// nans_slab_host.cpp — a C++ host that steps ONE world over several GPUs through the C ABI alone
// (include/nans_b200.h, nans_slab_*): no Python, no torch.  One process per GPU, as a multi-GPU build of the
// reference's host (code/sdl_nans.cpp) would be; the two 128-byte blobs of the set-up (the NCCL unique id and the
// CUDA-IPC descriptions of the arenas) travel over pipes between the processes -- any transport will do
// (INTEGRATION.md).  Per step nothing goes through the host: nans_slab_step queues the NCCL halo exchange and the
// cross-GPU dataflow solve on the world's stream.
//
//   nans_slab_host --gpus N [--side-x 16] [--ny 12] [--nz 16] [--steps 20] [--spacing 0.998]
//   nans_slab_host --single N ...        the same world (N slabs) on ONE GPU, stepped with nans_step
//
// Both print one line per slab: "slab r hash <fnv1a of pos/vel/ang/angvel of its bodies> contacts-on-rank <n>".
// The decomposed world is bit-identical to the single-GPU one, so the hashes must agree (tests/test_slab_gpu.py).
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/wait.h>
#include <unistd.h>

#include <vector>

#include "../../include/nans_b200.h"

#define CK(call)                                                                                              \
    do {                                                                                                      \
        int _rc = (call);                                                                                     \
        if (_rc) { fprintf(stderr, "nans_slab_host: %s failed (%d): %s\n", #call, _rc, nans_last_error()); exit(2); } \
    } while (0)

struct opts { int gpus, single, side_x, ny, nz, steps; float spacing; };

// deterministic jitter from the GLOBAL body index (so a rank can build its slab alone)
static float jitter(uint32_t gid, uint32_t axis)
{
    uint32_t h = gid * 2654435761u + axis * 40503u + 12345u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
    return ((float)(h & 0xffffff) / 16777216.0f - 0.5f) * 0.008f;
}

struct scene { int n; std::vector<float> pos, zeros, scale, mass, moi, radius; };

// slab `r` of the slab-major numbered pile (nans_projekat_b200/scenes.py: cube_pile_slabs), capacity rows
static scene build_slab(const opts &o, int r, int capacity)
{
    const int m = o.side_x * o.ny * o.nz;
    scene s;
    s.n = capacity;
    s.pos.assign(3 * (size_t)capacity, 0.f); s.zeros.assign(3 * (size_t)capacity, 0.f);
    s.scale.assign(3 * (size_t)capacity, 1.f); s.mass.assign(capacity, 1.f); s.moi.assign(capacity, 1.f);
    s.radius.assign(capacity, 0.f);
    for (int i = 0; i < capacity; ++i) { s.pos[3 * i + 1] = -1.0e6f; }          // ghost rows: parked, overwritten per step
    for (int i = 0; i < m; ++i) {
        const int y = i / (o.side_x * o.nz), rem = i % (o.side_x * o.nz), z = rem / o.side_x, x = rem % o.side_x;
        const uint32_t gid = (uint32_t)(r * m + i);
        s.pos[3 * i + 0] = (float)((double)(r * o.side_x + x) * o.spacing + 0.5) + jitter(gid, 0);
        s.pos[3 * i + 1] = (float)((double)y * o.spacing + 0.52) + jitter(gid, 1);
        s.pos[3 * i + 2] = (float)((double)z * o.spacing + 0.5) + jitter(gid, 2);
        s.moi[i] = (1.0f / 12.0f) * 2.0f;
    }
    return s;
}

static void upload(nans_world *w, const opts &o, const scene &s, int n_slabs)
{
    // one floor tile per slab, every rank holds all of them
    std::vector<float> sp(3 * n_slabs), sa(3 * n_slabs, 0.f), ss(3 * n_slabs), sm(n_slabs, 1.0e5f), si(n_slabs);
    const float wx = o.side_x * o.spacing, ez = o.nz * o.spacing;
    const float size_z = exp2f(ceilf(log2f(ez + 16.0f)));
    for (int r = 0; r < n_slabs; ++r) {
        sp[3 * r] = (r + 0.5f) * wx; sp[3 * r + 1] = -0.5f; sp[3 * r + 2] = ez / 2 + 0.07f;
        ss[3 * r] = wx; ss[3 * r + 1] = 1.0f; ss[3 * r + 2] = size_z;
        si[r] = (1.0e5f / 12.0f) * (2.0f * size_z * size_z);
    }
    nans_scene_view v;
    memset(&v, 0, sizeof(v));
    v.pos = const_cast<float *>(s.pos.data()); v.ang = const_cast<float *>(s.zeros.data());
    v.vel = const_cast<float *>(s.zeros.data()); v.angvel = const_cast<float *>(s.zeros.data());
    v.force = const_cast<float *>(s.zeros.data()); v.torque = const_cast<float *>(s.zeros.data());
    v.scale = const_cast<float *>(s.scale.data()); v.mass = const_cast<float *>(s.mass.data());
    v.moi = const_cast<float *>(s.moi.data()); v.radius = const_cast<float *>(s.radius.data());
    v.st_pos = sp.data(); v.st_ang = sa.data(); v.st_scale = ss.data(); v.st_mass = sm.data(); v.st_moi = si.data();
    CK(nans_world_upload(w, &v));
    CK(nans_rebuild_vertices(w));
}

static uint64_t hash_rows(nans_world *w, int capacity, int row0, int rows)
{
    std::vector<float> p(3 * (size_t)capacity), v(p.size()), a(p.size()), av(p.size());
    nans_scene_view d;
    memset(&d, 0, sizeof(d));
    d.pos = p.data(); d.vel = v.data(); d.ang = a.data(); d.angvel = av.data();
    CK(nans_world_download(w, &d));
    uint64_t h = 1469598103934665603ull;
    const std::vector<float> *f[4] = {&p, &v, &a, &av};
    for (int k = 0; k < 4; ++k) {
        const unsigned char *b = (const unsigned char *)(f[k]->data() + 3 * (size_t)row0);
        for (size_t i = 0; i < sizeof(float) * 3 * (size_t)rows; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    }
    return h;
}

static void xwrite(int fd, const void *p, size_t n) { if (write(fd, p, n) != (ssize_t)n) { perror("write"); exit(3); } }
static void xread(int fd, void *p, size_t n)
{
    size_t got = 0;
    while (got < n) { ssize_t k = read(fd, (char *)p + got, n - got); if (k <= 0) { perror("read"); exit(3); } got += (size_t)k; }
}

static int run_rank(const opts &o, int rank, int to_parent, int from_parent)
{
    const int m = o.side_x * o.ny * o.nz, halo_cap = 4 * o.ny * o.nz < 1024 ? 1024 : 4 * o.ny * o.nz, cap = m + halo_cap;
    nans_world_desc d;
    memset(&d, 0, sizeof(d));
    d.n_cubes = cap; d.n_statics = o.gpus; d.device = rank;
    d.max_pairs = 64 * cap; d.max_contacts = 24 * cap;          // the squeezed pile exceeds 8 contacts per body
    nans_world *w = nullptr;
    CK(nans_world_create(&d, &w));
    upload(w, o, build_slab(o, rank, cap), o.gpus);
    char id[128], blob[128];
    std::vector<char> blobs(128 * (size_t)o.gpus);
    if (rank == 0) { CK(nans_slab_unique_id(id)); xwrite(to_parent, id, 128); }
    xread(from_parent, id, 128);                                                  // broadcast
    CK(nans_slab_init(w, rank, o.gpus, id, m, rank * m, halo_cap));
    CK(nans_slab_ipc_handle(w, blob));
    xwrite(to_parent, blob, 128);
    xread(from_parent, blobs.data(), blobs.size());                               // all-gather
    CK(nans_slab_connect(w, blobs.data()));
    for (int k = 0; k < o.steps; ++k) CK(nans_slab_step(w, 1.0f / 60.0f));
    int32_t err = 0, live = 0;
    int64_t bytes = 0;
    CK(nans_slab_status(w, &err, &live, &bytes));
    nans_step_stats st;
    CK(nans_get_stats(w, &st));
    printf("slab %d hash %016llx contacts-on-rank %d ghosts %d\n", rank, (unsigned long long)hash_rows(w, cap, 0, m),
           st.n_contacts, live - m);
    fflush(stdout);
    nans_world_destroy(w);
    return 0;
}

static int run_single(const opts &o)
{
    const int m = o.side_x * o.ny * o.nz, n = o.single * m;
    nans_world_desc d;
    memset(&d, 0, sizeof(d));
    d.n_cubes = n; d.n_statics = o.single; d.device = 0;
    d.max_pairs = 64 * n; d.max_contacts = 24 * n;
    nans_world *w = nullptr;
    CK(nans_world_create(&d, &w));
    scene all;
    all.n = n;
    all.pos.resize(3 * (size_t)n); all.zeros.assign(3 * (size_t)n, 0.f); all.scale.assign(3 * (size_t)n, 1.f);
    all.mass.assign(n, 1.f); all.moi.assign(n, (1.0f / 12.0f) * 2.0f); all.radius.assign(n, 0.f);
    for (int r = 0; r < o.single; ++r) {
        const scene s = build_slab(o, r, m);
        memcpy(all.pos.data() + 3 * (size_t)r * m, s.pos.data(), sizeof(float) * 3 * (size_t)m);
    }
    upload(w, o, all, o.single);
    for (int k = 0; k < o.steps; ++k) CK(nans_step(w, 1.0f / 60.0f));
    nans_step_stats st;
    CK(nans_get_stats(w, &st));
    for (int r = 0; r < o.single; ++r)
        printf("slab %d hash %016llx contacts-in-world %d\n", r, (unsigned long long)hash_rows(w, n, r * m, m), st.n_contacts);
    nans_world_destroy(w);
    return 0;
}

int main(int argc, char **argv)
{
    opts o = {0, 0, 16, 12, 16, 20, 0.998f};
    for (int i = 1; i < argc; ++i) {
        auto val = [&](const char *name) { return !strcmp(argv[i], name) && i + 1 < argc; };
        if (val("--gpus")) o.gpus = atoi(argv[++i]);
        else if (val("--single")) o.single = atoi(argv[++i]);
        else if (val("--side-x")) o.side_x = atoi(argv[++i]);
        else if (val("--ny")) o.ny = atoi(argv[++i]);
        else if (val("--nz")) o.nz = atoi(argv[++i]);
        else if (val("--steps")) o.steps = atoi(argv[++i]);
        else if (val("--spacing")) o.spacing = (float)atof(argv[++i]);
        else { fprintf(stderr, "usage: %s --gpus N | --single N [--side-x X --ny Y --nz Z --steps S --spacing F]\n", argv[0]); return 1; }
    }
    if (o.single > 0) return run_single(o);
    if (o.gpus < 1) { fprintf(stderr, "--gpus N or --single N\n"); return 1; }
    // one process per GPU, forked BEFORE anything touches CUDA; the parent only relays the two blobs
    std::vector<int> up(o.gpus), down(o.gpus);
    std::vector<pid_t> pid(o.gpus);
    for (int r = 0; r < o.gpus; ++r) {
        int a[2], b[2];
        if (pipe(a) || pipe(b)) { perror("pipe"); return 3; }
        pid[r] = fork();
        if (pid[r] == 0) { close(a[0]); close(b[1]); return run_rank(o, r, a[1], b[0]); }
        close(a[1]); close(b[0]);
        up[r] = a[0]; down[r] = b[1];
    }
    char id[128];
    xread(up[0], id, 128);
    for (int r = 0; r < o.gpus; ++r) xwrite(down[r], id, 128);
    std::vector<char> blobs(128 * (size_t)o.gpus);
    for (int r = 0; r < o.gpus; ++r) xread(up[r], blobs.data() + 128 * (size_t)r, 128);
    for (int r = 0; r < o.gpus; ++r) xwrite(down[r], blobs.data(), blobs.size());
    int bad = 0;
    for (int r = 0; r < o.gpus; ++r) { int st = 0; waitpid(pid[r], &st, 0); bad |= !(WIFEXITED(st) && WEXITSTATUS(st) == 0); }
    return bad ? 4 : 0;
}
