/* ref_harness.c — TEST INFRASTRUCTURE (oracle/). Not part of the product path.
 *
 * Headless driver for the reference's own prebuilt plugin build/nans.so
 * (SURVEY.md §8c, Appendix A).  The reference sources cannot be compiled in this
 * image (GLEW/glm/SDL2 headers absent), but the shipped binary needs only
 * libstdc++/libm/libc plus 13 GL symbols, which are stubbed here.  The hot-path
 * functions are file-static in code/nans.cpp, so they are reached through the
 * ELF .symtab of the binary (read at load time, never hard-coded offsets):
 *
 *   Init                 code/nans.cpp:1551      IntegrateForces   :975
 *   DetectCollisions     code/nans.cpp:1352      SolveConstraints  :1539
 *   Constraint           code/nans.cpp:1021      IntegrateVelocities :1332
 *   CheckCollision       code/nans.cpp:907       UpdateVertices    :395
 *   FloorUpdateVertices  code/nans.cpp:380       SimUpdateAndRender :1719 (exported)
 *
 * Built into oracle/_ref/libnans_ref_harness.so by oracle/Makefile.  Only tests/,
 * __graft_entry__.smoke() and bench.py's CPU-baseline legs may load it.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <elf.h>
#include <fcntl.h>
#include <link.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "ref_layout.h"

/* ---- GL stubs.  nans.so imports nine GLEW function
 * pointer VARIABLES and four GL FUNCTIONS; all are no-ops here. ------------- */
static void stub_void() {}
static int stub_int0() { return 0; }
void (*__glewActiveTexture)() = stub_void;
void (*__glewBindVertexArray)() = stub_void;
int (*__glewGetUniformLocation)() = stub_int0;
void (*__glewUniform1f)() = stub_void;
void (*__glewUniform1i)() = stub_void;
void (*__glewUniform3f)() = stub_void;
void (*__glewUniform3fv)() = stub_void;
void (*__glewUniformMatrix4fv)() = stub_void;
void (*__glewUseProgram)() = stub_void;
void glBindTexture(unsigned a, unsigned b) { (void)a; (void)b; }
void glDrawArrays(unsigned a, int b, int c) { (void)a; (void)b; (void)c; }
void glDrawElements(unsigned a, int b, unsigned c, const void *d) { (void)a; (void)b; (void)c; (void)d; }
void glPolygonMode(unsigned a, unsigned b) { (void)a; (void)b; }

/* ---- function table --------------------------------------------------------- */
typedef void (*fn_state)(ref_sdl_state *);
typedef void (*fn_state_f)(ref_sdl_state *, float);
typedef void (*fn_state_i)(ref_sdl_state *, int);
typedef void (*fn_detect)(ref_sdl_state *, void *, void *, float, ref_stdvector *);
typedef void (*fn_solve)(ref_sdl_state *, float, ref_stdvector *);
typedef void (*fn_constraint)(ref_sdl_state *, ref_contact_pair *, float);
typedef int (*fn_check)(ref_sdl_state *, ref_contact_pair *, ref_stdvector *);
typedef void (*fn_entry)(ref_memory *, ref_sdl_input *, ref_sdl_render *, float);

static struct {
    void *handle;
    fn_state Init, FloorUpdateVertices;
    fn_state_f IntegrateForces, IntegrateVelocities;
    fn_state_i UpdateVertices;
    fn_detect DetectCollisions;
    fn_solve SolveConstraints;
    fn_constraint Constraint;
    fn_check CheckCollision;
    fn_entry SimUpdateAndRender;
} R;

static int sym_lookup(const unsigned char *img, size_t len, const char *name, uint64_t *out)
{
    const Elf64_Ehdr *eh = (const Elf64_Ehdr *)img;
    if (len < sizeof(*eh) || memcmp(eh->e_ident, ELFMAG, SELFMAG) != 0) return -1;
    const Elf64_Shdr *sh = (const Elf64_Shdr *)(img + eh->e_shoff);
    for (int i = 0; i < eh->e_shnum; ++i) {
        if (sh[i].sh_type != SHT_SYMTAB) continue;
        const Elf64_Sym *sym = (const Elf64_Sym *)(img + sh[i].sh_offset);
        size_t n = sh[i].sh_size / sizeof(Elf64_Sym);
        const char *str = (const char *)(img + sh[sh[i].sh_link].sh_offset);
        for (size_t k = 0; k < n; ++k)
            if (strcmp(str + sym[k].st_name, name) == 0) { *out = sym[k].st_value; return 0; }
    }
    return -1;
}

/* Promote this library to the global symbol scope so that nans.so's undefined
 * GL symbols bind to the stubs above even when we were loaded RTLD_LOCAL. */
static void promote_self(void)
{
    Dl_info info;
    if (dladdr((void *)&promote_self, &info) && info.dli_fname)
        (void)dlopen(info.dli_fname, RTLD_NOW | RTLD_GLOBAL | RTLD_NOLOAD);
}

int nansref_load(const char *so_path)
{
    if (R.handle) return 0;
    promote_self();
    int fd = open(so_path, O_RDONLY);
    if (fd < 0) return -1;
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return -1; }
    unsigned char *img = mmap(NULL, st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (img == MAP_FAILED) return -1;

    void *h = dlopen(so_path, RTLD_NOW | RTLD_LOCAL);
    if (!h) { fprintf(stderr, "nansref_load: %s\n", dlerror()); munmap(img, st.st_size); return -2; }
    void *entry = dlsym(h, "SimUpdateAndRender");
    uint64_t entry_off = 0;
    if (!entry || sym_lookup(img, st.st_size, "SimUpdateAndRender", &entry_off)) {
        munmap(img, st.st_size); dlclose(h); return -3;
    }
    char *base = (char *)entry - entry_off;
    int bad = 0;
#define RESOLVE(field, mangled)                                                        \
    do { uint64_t off;                                                                 \
         if (sym_lookup(img, st.st_size, mangled, &off)) { bad = 1;                    \
             fprintf(stderr, "nansref_load: missing symbol %s\n", mangled); }          \
         else *(void **)(&R.field) = base + off; } while (0)
    RESOLVE(Init, "_ZL4InitP9sdl_state");
    RESOLVE(FloorUpdateVertices, "_ZL19FloorUpdateVerticesP9sdl_state");
    RESOLVE(IntegrateForces, "_ZL15IntegrateForcesP9sdl_statef");
    RESOLVE(IntegrateVelocities, "_ZL19IntegrateVelocitiesP9sdl_statef");
    RESOLVE(UpdateVertices, "_ZL14UpdateVerticesP9sdl_statei");
    RESOLVE(DetectCollisions,
            "_ZL16DetectCollisionsP9sdl_stateP9sdl_inputP10sdl_renderfRSt6vectorI12contact_pairSaIS6_EE");
    RESOLVE(SolveConstraints, "_ZL16SolveConstraintsP9sdl_statefRSt6vectorI12contact_pairSaIS2_EE");
    RESOLVE(Constraint, "_ZL10ConstraintP9sdl_stateP12contact_pairf");
    RESOLVE(CheckCollision, "_ZL14CheckCollisionP9sdl_stateP12contact_pairRSt6vectorI6vertexSaIS4_EE");
#undef RESOLVE
    munmap(img, st.st_size);
    if (bad) { dlclose(h); return -4; }
    R.SimUpdateAndRender = (fn_entry)entry;
    R.handle = h;
    return 0;
}

int nansref_loaded(void) { return R.handle != NULL; }

/* ---- thin wrappers ---------------------------------------------------------- */
void nansref_init(ref_sdl_state *s) { R.Init(s); }
void nansref_integrate_forces(ref_sdl_state *s, float dt) { R.IntegrateForces(s, dt); }
void nansref_integrate_velocities(ref_sdl_state *s, float dt) { R.IntegrateVelocities(s, dt); }
void nansref_update_vertices(ref_sdl_state *s, int i) { R.UpdateVertices(s, i); }
void nansref_floor_update_vertices(ref_sdl_state *s) { R.FloorUpdateVertices(s); }
void nansref_constraint(ref_sdl_state *s, ref_contact_pair *p, float dt) { R.Constraint(s, p, dt); }
void nansref_sim_update_and_render(ref_memory *m, ref_sdl_input *in, ref_sdl_render *r, float dt)
{
    R.SimUpdateAndRender(m, in, r, dt);
}

/* std::vector blobs reused across calls (all-zero == empty; `end = begin` == clear()). */
static __thread ref_stdvector g_simplex;
static __thread ref_stdvector g_pairs;

int nansref_check_collision(ref_sdl_state *s, ref_contact_pair *p)
{
    return R.CheckCollision(s, p, &g_simplex);
}

/* DetectCollisions into a caller buffer; returns the number of contacts found
 * (may exceed cap, in which case only cap are copied). */
int nansref_detect_collisions(ref_sdl_state *s, float dt, ref_contact_pair *out, int cap)
{
    g_pairs.end = g_pairs.begin;
    R.DetectCollisions(s, NULL, NULL, dt, &g_pairs);
    int n = (int)((g_pairs.end - g_pairs.begin) / (long)sizeof(ref_contact_pair));
    int m = n < cap ? n : cap;
    if (m > 0) memcpy(out, g_pairs.begin, (size_t)m * sizeof(ref_contact_pair));
    return n;
}

/* SolveConstraints over a caller-provided list (the function only iterates). */
void nansref_solve_constraints(ref_sdl_state *s, float dt, ref_contact_pair *pairs, int n)
{
    ref_stdvector v;
    v.begin = (char *)pairs;
    v.end = v.begin + (size_t)n * sizeof(ref_contact_pair);
    v.cap = v.end;
    R.SolveConstraints(s, dt, &v);
}

/* The four physics stages exactly as code/nans.cpp:1758-1762 runs them.
 * Contacts of the step are copied to `out` (≤ cap); returns their count. */
int nansref_physics_step(ref_sdl_state *s, float dt, ref_contact_pair *out, int cap)
{
    g_pairs.end = g_pairs.begin;
    R.IntegrateForces(s, dt);
    R.DetectCollisions(s, NULL, NULL, dt, &g_pairs);
    R.SolveConstraints(s, dt, &g_pairs);
    R.IntegrateVelocities(s, dt);
    int n = (int)((g_pairs.end - g_pairs.begin) / (long)sizeof(ref_contact_pair));
    int m = n < cap ? n : cap;
    if (out && m > 0) memcpy(out, g_pairs.begin, (size_t)m * sizeof(ref_contact_pair));
    return n;
}

/* ---- batched narrowphase driver (config C3) --------------------------------
 * Shapes are described the way the new narrowphase takes them: kind 0 = box
 * given by its 8 world-space vertices (reference vertex order, code/nans.cpp:
 * 395-407), kind 1 = sphere (centre = pos, radius).  pair type follows
 * code/nans.h:71-87; for CF/SF shape B is the floor box.
 *   type   : [n] int32         posA,posB : [n][3]
 *   vertsA : [n][8][3] (ignored for spheres)    vertsB likewise
 *   radA, radB : [n]
 * outputs: hit [n] int32 ; N, PointA, PointB [n][3]
 */
void nansref_check_collision_batch(int n, const int32_t *type,
                                   const float *posA, const float *vertsA, const float *radA,
                                   const float *posB, const float *vertsB, const float *radB,
                                   int32_t *hit, float *outN, float *outPA, float *outPB)
{
    ref_sdl_state *s = calloc(1, sizeof(*s));
    s->CubeCount = 2;
    s->SphereCount = 2;
    for (int i = 0; i < n; ++i) {
        ref_contact_pair p;
        memset(&p, 0, sizeof(p));
        p.Type = type[i];
        int a_is_sphere = (type[i] == 3 || type[i] == 4);
        int b_is_sphere = (type[i] == 1 || type[i] == 3);
        int b_is_floor = (type[i] == 2 || type[i] == 4);
        p.IndexA = 0;
        p.IndexB = 1;
        if (a_is_sphere) {
            memcpy(&s->Spheres[0].Position, posA + 3 * i, 12);
            s->Spheres[0].Radius = radA[i];
        } else {
            memcpy(&s->Cubes[0].Position, posA + 3 * i, 12);
            memcpy(s->Cubes[0].Vertices, vertsA + 24 * i, 96);
        }
        if (b_is_floor) {
            memcpy(&s->Floor.Position, posB + 3 * i, 12);
            memcpy(s->Floor.Vertices, vertsB + 24 * i, 96);
        } else if (b_is_sphere) {
            memcpy(&s->Spheres[1].Position, posB + 3 * i, 12);
            s->Spheres[1].Radius = radB[i];
        } else {
            memcpy(&s->Cubes[1].Position, posB + 3 * i, 12);
            memcpy(s->Cubes[1].Vertices, vertsB + 24 * i, 96);
        }
        hit[i] = R.CheckCollision(s, &p, &g_simplex);
        memcpy(outN + 3 * i, &p.N, 12);
        memcpy(outPA + 3 * i, &p.PointA, 12);
        memcpy(outPB + 3 * i, &p.PointB, 12);
    }
    free(s);
}
