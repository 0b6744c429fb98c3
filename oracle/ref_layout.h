/* ref_layout.h — TEST INFRASTRUCTURE (oracle/). Not part of the product path.
 *
 * Plain-old-data mirrors of the structs the reference's prebuilt build/nans.so
 * was compiled with (code/nans.h:120-386).  The reference headers cannot be
 * included here (they pull in GLEW + glm, absent from this image), so the
 * layouts are re-declared from the DWARF of build/nans.so and
 * pinned by the static_asserts below.  glm::vec3 = 3 packed floats,
 * glm::mat4 = 16 floats column-major.
 */
#ifndef NANS_REF_LAYOUT_H
#define NANS_REF_LAYOUT_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
#define REF_STATIC_ASSERT(c, m) static_assert(c, m)
#else
#define REF_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif

typedef struct { float x, y, z; } ref_vec3;
typedef struct { float m[16]; } ref_mat4; /* column-major: m[4*col+row] */

/* code/nans.h:303-319 */
typedef struct {
    ref_mat4 Model;
    ref_vec3 Vertices[8];
    ref_vec3 Position, V, Forces;
    ref_vec3 Angles, W, Torque;
    float Size, Mass, MOI;
} ref_cube;
REF_STATIC_ASSERT(sizeof(ref_cube) == 244, "cube layout");
REF_STATIC_ASSERT(offsetof(ref_cube, Position) == 160, "cube.Position");
REF_STATIC_ASSERT(offsetof(ref_cube, Size) == 232, "cube.Size");

/* code/nans.h:321-336 */
typedef struct {
    ref_mat4 Model;
    ref_vec3 Position, V, Forces;
    ref_vec3 Angles, W, Torque;
    float Radius, Mass, MOI;
} ref_sphere;
REF_STATIC_ASSERT(sizeof(ref_sphere) == 148, "sphere layout");

/* code/nans.h:339-372 ; collision_type code/nans.h:71-87: CC=0 CS=1 CF=2 SS=3 SF=4 */
typedef struct {
    int32_t Type;
    float DLNormal, DLNormalSum;
    float DLTangent1, DLTangent1Sum;
    float DLTangent2, DLTangent2Sum;
    ref_vec3 PointA, PointB;
    ref_vec3 N, T1, T2;
    int32_t IndexA, IndexB;
} ref_contact_pair;
REF_STATIC_ASSERT(sizeof(ref_contact_pair) == 96, "contact_pair layout");
REF_STATIC_ASSERT(offsetof(ref_contact_pair, PointA) == 28, "contact_pair.PointA");
REF_STATIC_ASSERT(offsetof(ref_contact_pair, IndexA) == 88, "contact_pair.IndexA");

/* code/nans.h:279-293 */
typedef struct {
    float FOV, Pitch, Yaw, Speed;
    ref_vec3 Position, Target, Direction;
    ref_vec3 Up, Front, Right;
} ref_camera;
REF_STATIC_ASSERT(sizeof(ref_camera) == 88, "camera layout");

/* code/nans.h:96-108 */
typedef struct {
    ref_vec3 Position;
    float Size, Kc, Kl, Kq;
    ref_vec3 Ambient, Diffuse, Specular;
} ref_light;
REF_STATIC_ASSERT(sizeof(ref_light) == 64, "light layout");

/* libstdc++ std::vector<T>: {begin, end, end_of_storage}; all-zero == empty */
typedef struct { char *begin, *end, *cap; } ref_stdvector;

/* code/nans.h:374-386 */
typedef struct {
    ref_cube Cubes[16];
    ref_sphere Spheres[16];
    ref_cube Floor;
    ref_camera Camera;
    ref_light Lights[4];
    uint32_t CubeCount;
    uint32_t SphereCount;
    ref_stdvector Pairs;
} ref_sdl_state;
REF_STATIC_ASSERT(sizeof(ref_sdl_state) == 6896, "sdl_state layout");
REF_STATIC_ASSERT(offsetof(ref_sdl_state, Spheres) == 3904, "sdl_state.Spheres");
REF_STATIC_ASSERT(offsetof(ref_sdl_state, Floor) == 6272, "sdl_state.Floor");
REF_STATIC_ASSERT(offsetof(ref_sdl_state, Camera) == 6516, "sdl_state.Camera");
REF_STATIC_ASSERT(offsetof(ref_sdl_state, CubeCount) == 6860, "sdl_state.CubeCount");
REF_STATIC_ASSERT(offsetof(ref_sdl_state, Pairs) == 6872, "sdl_state.Pairs");

/* code/nans.h:120-165 */
typedef struct { int32_t HalfTransitionCount, EndedDown; } ref_button;
typedef struct {
    ref_button Buttons[13]; /* Fwd, Back, Left, Right, Shoot, DbgUp, DbgDown, DbgLeft,
                               DbgRight, DbgFwd, DbgBack, DbgReset, DbgContinue */
    float Sensitivity;
    int32_t XRel, YRel, X, Y;
} ref_sdl_input;
REF_STATIC_ASSERT(sizeof(ref_sdl_input) == 124, "sdl_input layout");

/* code/nans.h:208-225 */
typedef struct {
    uint32_t Shaders[3];
    uint32_t Textures[8];
    uint32_t VAOs[6];
    uint32_t VBOs[6];
    uint32_t *Indices;
    uint32_t *ModelIndices;
    uint32_t Num, ModelNum;
    uint32_t LightVAO;
    ref_mat4 View;
    ref_mat4 Projection;
} ref_sdl_render;
REF_STATIC_ASSERT(sizeof(ref_sdl_render) == 256, "sdl_render layout");
REF_STATIC_ASSERT(offsetof(ref_sdl_render, View) == 124, "sdl_render.View");

/* code/nans.h:227-235 */
typedef struct {
    void *PermanentStorage;
    uint64_t PermanentStorageSize;
    void *TransientStorage;
    uint64_t TransientStorageSize;
    int32_t IsInitialized;
} ref_memory;
REF_STATIC_ASSERT(sizeof(ref_memory) == 40, "memory layout");

/* code/nans.h:245-255 */
typedef struct { ref_vec3 P, SupA, SupB; } ref_vertex;
REF_STATIC_ASSERT(sizeof(ref_vertex) == 36, "vertex layout");

#endif
