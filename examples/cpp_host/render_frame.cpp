// A C++ host on the C ABI (include/ntracer_b200.h), the way the reference's own host would use it: the calls of
// INTEGRATION.md section A.3 -- scene create, camera, BlockingRenderer.render's ntr_render with the error mapping of
// PY_EXCEPT_HANDLERS (reference src/py_common.hpp:39-47), Scene.calculate_color's ntr_calculate_color -- for the
// scene of scripts/hypercube.py (a BoxScene seen from axis(2,-5), reference scripts/hypercube.py:309,356-361).
//
//   g++ -std=c++17 -Iinclude examples/cpp_host/render_frame.cpp -Lntracer_b200 -lntracer_b200 -o render_frame
//   LD_LIBRARY_PATH=ntracer_b200 ./render_frame 4 640 480 frame.rgb [n_gpus]
//
// Exit code 0 and the raw RGB8 frame in the file; exit code 3 with the library's message when there is no sm_100 device
// (the library has no CPU path).  tests/test_cpp_host.py builds and runs it.
#include <cstdio>
#include <cstdlib>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "ntracer_b200.h"

namespace {

// struct image_format (reference src/render.cpp:167-172) -> ntr_image_format
ntr_image_format rgb8(int width, int height) {
    ntr_image_format f{};
    f.width = width; f.height = height; f.pitch = width * 3;
    f.n_channels = 3;
    for (int c = 0; c < 3; ++c) {
        f.channels[c].f_r = c == 0; f.channels[c].f_g = c == 1; f.channels[c].f_b = c == 2; f.channels[c].f_c = 0;
        f.channels[c].bit_size = 8; f.channels[c].tfloat = 0;
    }
    f.bytes_per_pixel = 3; f.reversed = 0;
    return f;
}

// the exception mapping of the reference's bindings (PY_EXCEPT_HANDLERS): status -> C++ exception -> Python exception
struct aborted {};
void check(int rc) {
    switch (rc) {
        case NTR_OK: return;
        case NTR_ERR_ABORTED: throw aborted{};                                   // render() returns False
        case NTR_ERR_MEMORY: throw std::bad_alloc();                             // MemoryError
        case NTR_ERR_VALUE: throw std::invalid_argument(ntr_last_error());       // ValueError
        default: throw std::runtime_error(ntr_last_error());                     // RuntimeError ("already running", CUDA, no device)
    }
}

}  // namespace

int main(int argc, char **argv) {
    const int dim = argc > 1 ? atoi(argv[1]) : 4, width = argc > 2 ? atoi(argv[2]) : 640, height = argc > 3 ? atoi(argv[3]) : 480;
    const char *path = argc > 4 ? argv[4] : "frame.rgb";
    const int gpus = argc > 5 ? atoi(argv[5]) : 1;
    if (ntr_abi_version() != NTR_ABI_VERSION) { fprintf(stderr, "ABI mismatch\n"); return 2; }
    if (ntr_device_count() < 1) {
        ntr_scene_desc d{};
        d.dim = dim; d.kind = NTR_SCENE_BOX; d.fov = 0.8f;
        ntr_scene *sc = nullptr;
        const int rc = ntr_scene_create(&d, -1, &sc);
        fprintf(stderr, "no sm_100 device (status %d): %s\n", rc, ntr_last_error());
        return rc == NTR_ERR_NO_DEVICE ? 3 : 2;
    }
    try {
        ntr_scene_desc d{};
        d.dim = dim; d.kind = NTR_SCENE_BOX; d.batch_size = 1; d.root = NTR_NULL_NODE; d.fov = 0.8f;      // box_scene: fov only
        std::vector<float> origin(dim, 0.0f), axes((size_t)dim * dim, 0.0f);
        origin[2] = -5.0f;                                                       // cam.translate(Vector.axis(2,-5))
        for (int i = 0; i < dim; ++i) axes[(size_t)i * dim + i] = 1.0f;          // camera(): identity orientation
        const ntr_image_format fmt = rgb8(width, height);
        std::vector<unsigned char> frame((size_t)fmt.pitch * height);
        float centre[3] = {0, 0, 0};
        unsigned long long launches = 0;
        if (gpus > 1) {
            ntr_group *g = nullptr;
            check(ntr_group_create(&d, gpus, nullptr, &g));
            check(ntr_group_set_camera(g, origin.data(), axes.data()));
            check(ntr_group_render(g, &fmt, frame.data(), frame.size()));
            launches = ntr_group_launch_count(g);
            ntr_group_destroy(g);
        } else {
            ntr_scene *sc = nullptr;
            check(ntr_scene_create(&d, -1, &sc));
            check(ntr_scene_set_camera(sc, origin.data(), axes.data()));
            check(ntr_render(sc, &fmt, frame.data(), frame.size()));             // obj_BlockingRenderer_render
            check(ntr_calculate_color(sc, width / 2, height / 2, width, height, centre));   // obj_Scene_calculate_color
            launches = ntr_launch_count(sc);
            // a destination that is too small is the caller's error, reported like im_check_buffer_size does
            if (ntr_render(sc, &fmt, frame.data(), 16) != NTR_ERR_VALUE) { fprintf(stderr, "short buffer accepted\n"); return 2; }
            ntr_scene_destroy(sc);
        }
        FILE *f = fopen(path, "wb");
        if (!f || fwrite(frame.data(), 1, frame.size(), f) != frame.size()) { fprintf(stderr, "cannot write %s\n", path); return 2; }
        fclose(f);
        printf("rendered %dx%d %d-D BoxScene on %d GPU(s), %llu kernel launch(es), centre colour %.7g %.7g %.7g\n", width, height, dim,
               gpus, launches, centre[0], centre[1], centre[2]);
    } catch (const aborted &) {
        printf("aborted\n");
    } catch (const std::exception &e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 2;
    }
    return 0;
}
