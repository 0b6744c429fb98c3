/* nans_oracle.h — TEST INFRASTRUCTURE (oracle/). Not part of the product path.
 *
 * CPU restatement, in plain C, of the reference's rigid-body step
 * (IntegrateForces -> DetectCollisions(GJK+EPA) -> SolveConstraints ->
 * IntegrateVelocities -> model/vertex rebuild; code/nans.cpp:1758-1762 and
 * :1870-1881,1913-1941), generalised from the reference's 16+16+1 bodies to
 * arbitrary counts.  Parity is PINNED: tests/test_oracle_vs_ref.py proves every
 * function here bit-identical to the reference's own prebuilt build/nans.so over
 * random states (see oracle/README.md for the counts), and tests/golden/ holds
 * nans.so outputs for the boxes where /root/reference is absent.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may use it.
 */
#ifndef NANS_ORACLE_H
#define NANS_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* collision_type, code/nans.h:71-87 */
enum { ORC_CC = 0, ORC_CS = 1, ORC_CF = 2, ORC_SS = 3, ORC_SF = 4 };

/* World: dynamic bodies are cubes [0,n_cubes) followed by spheres
 * [n_cubes, n_cubes+n_spheres) in every per-body array ("nb" rows).  Statics are
 * floor-type slabs (the reference has exactly one, code/nans.cpp:1667-1678):
 * boxes that enter JMJ with their mass/MOI but never receive impulses. */
typedef struct {
    int32_t n_cubes, n_spheres, n_statics;
    float *pos, *vel, *force;      /* [nb][3] */
    float *ang, *angvel, *torque;  /* [nb][3]  (Euler "Angles", code/nans.h:313) */
    float *mass, *moi;             /* [nb] */
    float *scale;                  /* [nb][3] model scale (cubes); unused for spheres */
    float *radius;                 /* [nb]    (spheres) */
    float *verts;                  /* [n_cubes][8][3] world-space collision vertices */
    float *st_pos, *st_ang, *st_scale; /* [n_statics][3] */
    float *st_mass, *st_moi;       /* [n_statics] */
    float *st_verts;               /* [n_statics][8][3] */
} oracle_world;

/* One contact, the meaningful fields of contact_pair (code/nans.h:339-372).
 * a/b index cubes for CC; cube,sphere for CS; cube,static for CF; sphere,sphere
 * for SS; sphere,static for SF (sphere indices are 0-based within spheres). */
typedef struct {
    int32_t type, a, b;
    float point_a[3], point_b[3], n[3];
} oracle_contact;

/* shape for the stand-alone narrowphase entry: kind 0 = box (8 world verts), 1 = sphere */
typedef struct {
    int32_t kind;
    float pos[3];       /* body centre (GJK start direction, sphere support) */
    float radius;
    float verts[24];
} oracle_shape;

typedef struct {
    int32_t gjk_result;   /* evolve_result after the GJK loop: 0 none, 1 found, 2 still evolving */
    int32_t gjk_iters;
    int32_t epa_iters;    /* 0 if EPA not entered */
    int32_t max_faces, max_edges; /* EPA high-water marks (capacity planning for the device arenas) */
    int32_t emptied;      /* EPA iterations that left the triangle list EMPTY (the next one reads the stale Triangle[0]) */
} oracle_np_stats;

/* code/nans.cpp:907-966 (+ :572-769 GJK, :788-904 EPA). Returns the bool32 result. */
int oracle_check_collision(const oracle_shape *A, const oracle_shape *B,
                           float outN[3], float outPA[3], float outPB[3], oracle_np_stats *stats);

void oracle_check_collision_batch(int n, const int32_t *type,
                                  const float *posA, const float *vertsA, const float *radA,
                                  const float *posB, const float *vertsB, const float *radB,
                                  int32_t *hit, int32_t *gjk, float *outN, float *outPA, float *outPB,
                                  oracle_np_stats *stats_or_null);

/* code/nans.cpp:51-78, :975-1018 (vertex refresh is a no-op here: verts are state) */
void oracle_integrate_forces(oracle_world *w, float dt);
/* code/nans.cpp:1352-1536; all pairs in reference order. prefilter!=0 skips pairs whose
 * inflated AABBs are disjoint (NOT the reference algorithm; same result, see tests). */
int oracle_detect_collisions(const oracle_world *w, oracle_contact *out, int cap, int prefilter);
/* code/nans.cpp:1021-1329 applied to one contact */
void oracle_constraint(oracle_world *w, const oracle_contact *c, float dt);
/* code/nans.cpp:1539-1548 */
void oracle_solve_constraints(oracle_world *w, float dt, const oracle_contact *c, int n);
/* code/nans.cpp:1332-1349 */
void oracle_integrate_velocities(oracle_world *w, float dt);
/* code/nans.cpp:1870-1881,1913-1941 + :380-407 : Model = T*Rx*Ry*Rz*S -> 8 vertices */
void oracle_rebuild_vertices(oracle_world *w);
void oracle_model_vertices(const float pos[3], const float ang[3], const float scale[3],
                           float model_out[16], float verts_out[24]);
/* whole step: forces, detect, solve, velocities, rebuild. returns contact count */
int oracle_step(oracle_world *w, float dt, oracle_contact *scratch, int cap, int prefilter);

#ifdef __cplusplus
}
#endif
#endif
