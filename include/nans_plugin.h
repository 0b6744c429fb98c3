/* nans_plugin.h — the game-layer plugin boundary, layout-compatible with the reference.
 *
 * The reference host (code/sdl_nans.cpp:395-431, 913, 922-930, 986) dlopen()s "nans.so", resolves
 * the single C symbol below, and calls it once per frame with host-owned memory:
 *
 *     extern "C" void SimUpdateAndRender(memory *Memory, sdl_input *Input, sdl_render *Render, real32 dt)
 *     (code/nans.h:396-397; definition code/nans.cpp:1719)
 *
 * The three structs are re-declared here (the reference's header cannot be included: it pulls in
 * GLEW and glm) with the exact layout of the reference build — sizes/offsets are those of the
 * DWARF in build/nans.so and are pinned by static_asserts.  `sdl_state` is private
 * to the plugin (the host never reads it), so the new plugin keeps its own state there instead:
 * a host mirror of the world plus the handle of the device world (see nans_plugin.cpp).
 */
#ifndef NANS_PLUGIN_H
#define NANS_PLUGIN_H
#include <stddef.h>
#include <stdint.h>

typedef int32_t bool32;
typedef float real32;

/* code/nans.h:227-235 */
typedef struct memory {
    void *PermanentStorage;
    uint64_t PermanentStorageSize;
    void *TransientStorage;
    uint64_t TransientStorageSize;
    bool32 IsInitialized;
} memory;

/* code/nans.h:129-133 */
typedef struct sdl_button_state {
    int32_t HalfTransitionCount;
    bool32 EndedDown;
} sdl_button_state;

/* code/nans.h:135-159 + 120-127: keyboard (13 buttons) then mouse */
typedef struct sdl_input {
    union {
        sdl_button_state Buttons[13];
        struct {
            sdl_button_state MoveForward, MoveBack, MoveLeft, MoveRight, ShootAction;
            sdl_button_state DebugUp, DebugDown, DebugLeft, DebugRight, DebugForward, DebugBack;
            sdl_button_state DebugReset, DebugContinue;
        };
    } KeyboardController;
    struct {
        real32 Sensitivity;
        int32_t XRel, YRel, X, Y;
    } MouseController;
} sdl_input;

/* code/nans.h:208-225; glm::mat4 = 16 floats, column-major */
typedef struct sdl_render {
    uint32_t Shaders[3];
    uint32_t Textures[8];
    uint32_t VAOs[6];
    uint32_t VBOs[6];
    uint32_t *Indices;
    uint32_t *ModelIndices;
    uint32_t Num;
    uint32_t ModelNum;
    uint32_t LightVAO;
    float View[16];
    float Projection[16];
} sdl_render;

#ifdef __cplusplus
static_assert(sizeof(memory) == 40, "memory layout");
static_assert(sizeof(sdl_input) == 124 && offsetof(sdl_input, MouseController) == 104, "sdl_input layout");
static_assert(sizeof(sdl_render) == 256 && offsetof(sdl_render, View) == 124, "sdl_render layout");
#endif

#define SIM_UPDATE_AND_RENDER(name) void name(memory *Memory, sdl_input *Input, sdl_render *Render, real32 dt)
typedef SIM_UPDATE_AND_RENDER(sim_update_and_render);

#ifdef __cplusplus
extern "C" {
#endif
/* The plugin's one export. */
SIM_UPDATE_AND_RENDER(SimUpdateAndRender);

/* Read-only peek at the plugin's host mirror (headless hosts / tests; the reference host does not
 * need it).  Returns body count; pointers (valid until the next frame) to [nb][3] arrays. */
typedef struct nans_plugin_view {
    int32_t n_cubes, n_spheres, n_contacts, frame;
    const float *pos, *ang, *vel, *angvel;
    float cam_pos[3], cam_front[3], cam_yaw, cam_pitch;
} nans_plugin_view;
int NansPluginPeek(const memory *Memory, nans_plugin_view *out);
#ifdef __cplusplus
}
#endif
#endif
