// nans_host.cpp — headless host executable `nans`: the platform layer of the reference
// (code/sdl_nans.cpp) without SDL/OpenGL.  It keeps exactly the parts that touch the plugin
// boundary:
//   * one zero-filled mmap of 64 MiB permanent + 256 MiB transient storage  (:541-555)
//   * dlopen("nans.so") + dlsym("SimUpdateAndRender"), no-op stub on failure (:395-417)
//   * hot reload: stat() the plugin every frame, on a new mtime dlclose -> sleep 100 ms -> dlopen
//     (:922-930); all simulation state survives because it lives in the host's memory block
//   * double-buffered input carried across frames (:932-952), dt = 0 on the first frame and the
//     measured frame time afterwards (:902,999) unless --dt pins it, optional 60 Hz pacing (:964-982)
// plus scripted input and a trajectory dump for tests and benchmarks.
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <string>

#include "../../include/nans_plugin.h"

typedef int (*peek_fn)(const memory *, nans_plugin_view *);

static SIM_UPDATE_AND_RENDER(SimUpdateAndRenderStub) { (void)Memory; (void)Input; (void)Render; (void)dt; }

struct sim_code {           // sdl_sim_code, code/sdl_nans.h:6-13
    void *handle;
    time_t last_write;
    sim_update_and_render *UpdateAndRender;
    peek_fn Peek;
    bool valid;
};

static time_t last_write_time(const char *path)
{
    struct stat st;
    return stat(path, &st) == 0 ? st.st_mtime : 0;
}

static sim_code load_sim_code(const char *path)    // SDLLoadSimCode, code/sdl_nans.cpp:396-417
{
    sim_code c;
    memset(&c, 0, sizeof(c));
    c.last_write = last_write_time(path);
    c.handle = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (c.handle) {
        c.UpdateAndRender = (sim_update_and_render *)dlsym(c.handle, "SimUpdateAndRender");
        c.Peek = (peek_fn)dlsym(c.handle, "NansPluginPeek");
        c.valid = c.UpdateAndRender != NULL;
    } else {
        fprintf(stderr, "nans host: dlopen(%s): %s\n", path, dlerror());
    }
    if (!c.valid) c.UpdateAndRender = SimUpdateAndRenderStub;
    return c;
}

static void unload_sim_code(sim_code *c)           // SDLUnloadSimCode, code/sdl_nans.cpp:420-431
{
    if (c->handle) dlclose(c->handle);
    c->handle = NULL;
    c->valid = false;
    c->UpdateAndRender = SimUpdateAndRenderStub;
    c->Peek = NULL;
}

static double now_s()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int main(int argc, char **argv)
{
    std::string plugin;
    {   // default: nans.so next to this executable (the reference links with -rpath '$ORIGIN')
        char self[4096];
        ssize_t n = readlink("/proc/self/exe", self, sizeof(self) - 1);
        self[n > 0 ? n : 0] = 0;
        plugin = self;
        plugin = plugin.substr(0, plugin.find_last_of('/') + 1) + "nans.so";
    }
    int frames = 1000, reload_at = -1;
    float fixed_dt = -1.0f;
    bool pace = false, script_demo = false, quiet = false;
    const char *dump_path = NULL;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--plugin") && i + 1 < argc) plugin = argv[++i];
        else if (!strcmp(argv[i], "--frames") && i + 1 < argc) frames = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--dt") && i + 1 < argc) fixed_dt = (float)atof(argv[++i]);
        else if (!strcmp(argv[i], "--pace")) pace = true;
        else if (!strcmp(argv[i], "--script") && i + 1 < argc) script_demo = !strcmp(argv[++i], "demo");
        else if (!strcmp(argv[i], "--dump") && i + 1 < argc) dump_path = argv[++i];
        else if (!strcmp(argv[i], "--reload-at") && i + 1 < argc) reload_at = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--quiet")) quiet = true;
        else {
            fprintf(stderr, "usage: nans [--plugin nans.so] [--frames N] [--dt S] [--pace] [--script demo] "
                            "[--dump FILE] [--reload-at FRAME] [--quiet]\n");
            return 2;
        }
    }

    // memory, code/sdl_nans.cpp:541-555
    memory SimMemory;
    memset(&SimMemory, 0, sizeof(SimMemory));
    SimMemory.PermanentStorageSize = 64ull << 20;
    SimMemory.TransientStorageSize = 256ull << 20;
    const size_t total = SimMemory.PermanentStorageSize + SimMemory.TransientStorageSize;
    void *block = mmap(NULL, total, PROT_READ | PROT_WRITE, MAP_ANONYMOUS | MAP_PRIVATE, -1, 0);
    if (block == MAP_FAILED) { perror("mmap"); return 1; }
    SimMemory.PermanentStorage = block;
    SimMemory.TransientStorage = (char *)block + SimMemory.PermanentStorageSize;

    sim_code Sim = load_sim_code(plugin.c_str());
    if (!Sim.valid) fprintf(stderr, "nans host: running the no-op stub (plugin not loaded)\n");

    sdl_input Input[2];
    memset(Input, 0, sizeof(Input));
    sdl_input *NewInput = &Input[0], *OldInput = &Input[1];
    sdl_render Render;
    memset(&Render, 0, sizeof(Render));

    FILE *dump = dump_path ? fopen(dump_path, "wb") : NULL;
    float dt = 0.0f;                      // first frame, code/sdl_nans.cpp:902
    const double target = 1.0 / 60.0;
    double t_last = now_s(), t_sim = 0.0;
    int reloads = 0;
    for (int frame = 0; frame < frames; ++frame) {
        // hot reload, code/sdl_nans.cpp:922-930 (--reload-at forces one, as if the file had been rebuilt)
        const time_t wt = last_write_time(plugin.c_str());
        if ((wt != 0 && wt != Sim.last_write) || frame == reload_at) {
            unload_sim_code(&Sim);
            usleep(100 * 1000);
            Sim = load_sim_code(plugin.c_str());
            ++reloads;
        }
        // input double buffering, code/sdl_nans.cpp:932-952
        memset(NewInput, 0, sizeof(*NewInput));
        for (int b = 0; b < 13; ++b)
            NewInput->KeyboardController.Buttons[b].EndedDown = OldInput->KeyboardController.Buttons[b].EndedDown;
        NewInput->MouseController.Sensitivity = 0.5f;
        NewInput->MouseController.X = OldInput->MouseController.X;
        NewInput->MouseController.Y = OldInput->MouseController.Y;
        if (script_demo) {   // aim at frame 200, one-frame shot at 201
            NewInput->KeyboardController.ShootAction.EndedDown = (frame == 201);
            if (frame == 200) { NewInput->MouseController.XRel = 127; NewInput->MouseController.YRel = -44; }
        }
        if (pace) {          // code/sdl_nans.cpp:964-982
            double el = now_s() - t_last;
            if (el < target) usleep((useconds_t)((target - el) * 1e6));
        }
        const double t0 = now_s();
        Sim.UpdateAndRender(&SimMemory, NewInput, &Render, dt);
        const double t1 = now_s();
        t_sim += t1 - t0;
        dt = fixed_dt >= 0.0f ? fixed_dt : (float)(t1 - t_last);   // code/sdl_nans.cpp:999
        t_last = t1;
        if (dump && Sim.Peek) {
            nans_plugin_view v;
            const int nb = Sim.Peek(&SimMemory, &v);
            if (nb > 0) {
                int32_t hdr[4] = {frame, nb, v.n_contacts, 0};
                fwrite(hdr, sizeof(hdr), 1, dump);
                fwrite(v.pos, 12, nb, dump); fwrite(v.ang, 12, nb, dump);
                fwrite(v.vel, 12, nb, dump); fwrite(v.angvel, 12, nb, dump);
                float cam[8] = {v.cam_pos[0], v.cam_pos[1], v.cam_pos[2], v.cam_front[0], v.cam_front[1],
                                v.cam_front[2], v.cam_yaw, v.cam_pitch};
                fwrite(cam, sizeof(cam), 1, dump);
                fwrite(Render.View, 64, 1, dump); fwrite(Render.Projection, 64, 1, dump);
            }
        }
        sdl_input *tmp = NewInput; NewInput = OldInput; OldInput = tmp;   // code/sdl_nans.cpp:1009-1011
    }
    if (dump) fclose(dump);
    if (!quiet) {
        nans_plugin_view v;
        int nb = Sim.Peek ? Sim.Peek(&SimMemory, &v) : 0;
        printf("frames %d  sim %.3f ms/frame  reloads %d  bodies %d  contacts %d\n", frames,
               1e3 * t_sim / (frames > 0 ? frames : 1), reloads, nb, nb > 0 ? v.n_contacts : 0);
        for (int i = 0; i < nb && i < 8; ++i)
            printf("body %d pos %.6f %.6f %.6f\n", i, v.pos[3 * i], v.pos[3 * i + 1], v.pos[3 * i + 2]);
    }
    unload_sim_code(&Sim);
    munmap(block, total);
    return 0;
}
