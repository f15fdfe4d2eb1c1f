// splat_demo -- the reference's viewer loop (src/main.rs:9-30, :69-79; bin/02_ply_demo.rs) without the
// window, written against include/splat_pipeline.hpp: load a scene, orbit the camera, clear the colour
// buffer, render_to_buffer, write the frames out.  It is also what tests/test_cpp_host.py drives:
// every mode dumps exactly what it computed or handed to the library, in raw little-endian arrays.
//
//   splat_demo camera  H W x y z yaw pitch OUT        splat_camera struct bytes of that pose -> OUT
//   splat_demo ply     FILE OUT                       load_from_ply -> from_vec: N (u64) then positions |
//                                                     scales | opacities | rotations | sh -> OUT
//   splat_demo trim    SRC DST COUNT                  the first COUNT vertices of a PLY (00_ply_load.rs:9-63)
//   splat_demo naive   OUT                            the 4-Gaussian scene, same dump
//   splat_demo render  FILE|naive H W x y z FRAMES YAW_STEP PIPELINE(1|2) CLEARED(0|1) OUT
//                                                     FRAMES frames, the camera yawed by YAW_STEP before each;
//                                                     per frame: splat_camera bytes then W*H u32 pixels -> OUT
//                                                     env SPLAT_DEMO_DEVICES=0,1,..: a group context over those GPUs
// Exit codes: 0 ok, 2 usage, 3 a splat_b200::Error (message on stderr).  There is no CPU path: `render`
// on a box without a usable GPU exits 3 with the library's message.
#include <cstdio>
#include <cstdlib>
#include <string>

#include "splat_pipeline.hpp"

using namespace splat_b200;

static void put(std::FILE *f, const void *p, size_t bytes) {
  if (bytes && std::fwrite(p, 1, bytes, f) != bytes) throw Error(0, "short write");
}

static void dump_list(const GaussianList &l, const char *path) {
  std::FILE *f = std::fopen(path, "wb");
  if (!f) throw Error(0, std::string("cannot write ") + path);
  const uint64_t n = l.num_gaussians;
  put(f, &n, sizeof n);
  put(f, l.positions.data(), l.positions.size() * 4);
  put(f, l.scales.data(), l.scales.size() * 4);
  put(f, l.opacities.data(), l.opacities.size() * 4);
  put(f, l.rotations.data(), l.rotations.size() * 4);
  put(f, l.sh.data(), l.sh.size() * 4);
  std::fclose(f);
}

static int usage() {
  std::fprintf(stderr, "usage: splat_demo camera|ply|trim|naive|render ...  (see the header of splat_demo.cpp)\n");
  return 2;
}

int main(int argc, char **argv) {
  try {
    if (argc < 2) return usage();
    const std::string mode = argv[1];
    if (mode == "camera" && argc == 10) {
      Camera cam((float)std::atof(argv[2]), (float)std::atof(argv[3]),
                 Vec3{(float)std::atof(argv[4]), (float)std::atof(argv[5]), (float)std::atof(argv[6])});
      cam.update_yaw_angle((float)std::atof(argv[7]));
      cam.update_pitch_angle((float)std::atof(argv[8]));
      cam.update_camera_pose();
      const splat_camera s = camera_struct(cam);
      std::FILE *f = std::fopen(argv[9], "wb");
      if (!f) throw Error(0, "cannot write output");
      put(f, &s, sizeof s);
      std::fclose(f);
      return 0;
    }
    if (mode == "ply" && argc == 4) {
      dump_list(GaussianList::from_vec(load_from_ply(argv[2])), argv[3]);
      return 0;
    }
    if (mode == "trim" && argc == 5) {
      std::printf("%zu\n", trim_ply(argv[2], argv[3], (size_t)std::atoll(argv[4])));
      return 0;
    }
    if (mode == "naive" && argc == 3) {
      dump_list(GaussianList::naive_gaussians(), argv[2]);
      return 0;
    }
    if (mode == "render" && argc == 13) {
      const std::string scene = argv[2];
      const float H = (float)std::atof(argv[3]), W = (float)std::atof(argv[4]);
      const Vec3 start{(float)std::atof(argv[5]), (float)std::atof(argv[6]), (float)std::atof(argv[7])};
      const int frames = std::atoi(argv[8]);
      const float yaw_step = (float)std::atof(argv[9]);
      const int which = std::atoi(argv[10]);
      const bool cleared = std::atoi(argv[11]) != 0;
      std::vector<Gaussian> gs = scene == "naive" ? naive_gaussians() : load_from_ply(scene);
      Camera camera(H, W, start);                                  // main.rs:21 / 02_ply_demo.rs:22
      std::FILE *f = std::fopen(argv[12], "wb");
      if (!f) throw Error(0, "cannot write output");
      Buffer2d<uint32_t> color = Buffer2d<uint32_t>::fill({(size_t)W, (size_t)H}, 0u);   // main.rs:28
      auto loop = [&](auto &pipeline) {
        for (int i = 0; i < frames; ++i) {
          pipeline.camera.update_yaw_angle(yaw_step);              // what the arrow keys do, main.rs:50-66
          pipeline.camera.update_camera_pose();                    // main.rs:70
          const splat_camera s = camera_struct(pipeline.camera);
          put(f, &s, sizeof s);
          if constexpr (std::is_same_v<std::decay_t<decltype(pipeline)>, GaussianSplatPipeline02>) {
            if (cleared) {
              pipeline.render_cleared_to_buffer(color, 0u);
              put(f, color.raw(), (size_t)W * (size_t)H * 4);
              continue;
            }
          }
          color.fill(0u);                                          // main.rs:73
          pipeline.render_to_buffer(color);                        // main.rs:74
          put(f, color.raw(), (size_t)W * (size_t)H * 4);
        }
      };
      std::vector<int32_t> devices;
      if (const char *d = std::getenv("SPLAT_DEMO_DEVICES")) {
        for (const char *q = d; *q;) {
          devices.push_back((int32_t)std::strtol(q, const_cast<char **>(&q), 10));
          if (*q == ',') ++q;
        }
      }
      if (devices.empty()) devices.push_back(0);
      if (which == 1) {
        GaussianSplatPipeline01 p(std::move(gs), camera, devices);
        loop(p);
      } else {
        GaussianSplatPipeline02 p(GaussianList::from_vec(gs), camera, devices);
        loop(p);
        const splat_timings t = p.timings();
        std::fprintf(stderr, "splat_demo: %llu Gaussians, %llu visible, %llu tile instances, last frame %.3f ms\n",
                     (unsigned long long)t.n_gaussians, (unsigned long long)t.n_visible, (unsigned long long)t.n_instances, t.total_ms);
      }
      std::fclose(f);
      return 0;
    }
    return usage();
  } catch (const Error &e) {
    std::fprintf(stderr, "splat_demo: error %d: %s\n", e.code, e.what());
    return 3;
  }
}
