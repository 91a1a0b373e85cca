// Drives barbu::Hair (include/barbu_hair.hpp) the way the reference's Renderer drives its Hair module
// (core/renderer.cc:17-23,69-81; Application.cc:38-39): init, setup(scalp), per frame set_bounding_sphere + update(dt).
// Usage: hair_adaptor_main <in.bin> <out.bin> [devices, e.g. 0,1 or 0,0 — sharded over those CUDA devices (bh_group_*)]
//   in : int64 S, int64 F, int32 N, int32 nframes, uint32 seed, float dt, float scale, float sphere[4], int32 math,
//        float pos[S*3], float nrm[S*3], int32 tri[F*3]
//   out: int64 V, int64 nelems, float pos4[V*4], float vel4[V*4], float tan4[V*4], int32 patch[nelems],
//        int64 nstream, float stream4[nstream*4]   (tess-stream of the final state, default tessellation; a save_state /
//        load_state round trip is made in between and must not change it)
// Exit codes: 0 ok, 2 usage/io, 3 module not initialised after setup.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/barbu_hair.hpp"

int main(int argc, char** argv) {
  if (argc != 3 && argc != 4) return 2;
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  std::int64_t S = 0, F = 0; std::int32_t N = 0, nframes = 0, math = 0; std::uint32_t seed = 0; float dt = 0, scale = 0, sphere[4];
  bool ok = std::fread(&S, 8, 1, f) == 1 && std::fread(&F, 8, 1, f) == 1 && std::fread(&N, 4, 1, f) == 1 && std::fread(&nframes, 4, 1, f) == 1 &&
            std::fread(&seed, 4, 1, f) == 1 && std::fread(&dt, 4, 1, f) == 1 && std::fread(&scale, 4, 1, f) == 1 &&
            std::fread(sphere, 4, 4, f) == 4 && std::fread(&math, 4, 1, f) == 1;
  std::vector<float> pos(ok ? 3 * S : 0), nrm(ok ? 3 * S : 0);
  std::vector<std::int32_t> tri(ok ? 3 * F : 0);
  ok = ok && std::fread(pos.data(), 4, pos.size(), f) == pos.size() && std::fread(nrm.data(), 4, nrm.size(), f) == nrm.size() &&
       std::fread(tri.data(), 4, tri.size(), f) == tri.size();
  std::fclose(f);
  if (!ok) return 2;

  barbu::Hair hair;
  hair.init();
  hair.update(dt);                                   // before setup: must be a silent no-op (hair.cc:90-93)
  if (hair.initialized()) return 3;
  hair.setup(barbu::ScalpMesh{});                    // missing scalp: logs, stays uninitialised (hair.cc:45-48)
  if (hair.initialized()) return 3;

  hair.params().b200.ncontrol_points = N;
  hair.params().b200.seed = seed;
  hair.params().b200.math = math;
  hair.params().render.lengthScale = scale;
  if (argc == 4)
    for (const char* c = argv[3]; *c;) { hair.params().b200.devices.push_back(static_cast<int>(std::strtol(c, const_cast<char**>(&c), 10))); if (*c == ',') ++c; }
  const bool sharded = hair.params().b200.devices.size() > 1;
  barbu::ScalpMesh scalp;
  scalp.positions = pos.data(); scalp.normals = nrm.data(); scalp.nvertices = S; scalp.indices = tri.data(); scalp.nfaces = F;
  hair.setup(scalp);
  if (!hair.initialized()) return 3;
  for (int frame = 0; frame < nframes; ++frame) {    // Renderer::update: collider feed, then the step
    hair.set_bounding_sphere(sphere);
    hair.update(dt);
  }
  const std::int64_t V = hair.nvertices();
  std::vector<float> p4(4 * V), v4(4 * V), t4(4 * V);
  if (!hair.download(p4.data(), v4.data(), t4.data())) return 3;
  const std::int64_t nelems = static_cast<std::int64_t>(hair.patch_indices().size());
  // the render-side half: tess-stream, then state file round trip, then the stream again — identical
  if (sharded) {                                     // the gathered position plane on the render GPU must be the downloaded one
    float ms = -1.f;
    if (!hair.gather_positions(&ms) || ms < 0.f) return 3;
  }
  std::vector<float> stream4(sharded ? 0 : 4 * static_cast<size_t>(hair.stream_count())), stream4b(stream4.size());
  const std::int64_t nstream = sharded ? 0 : hair.stream(stream4.data());
  if (!sharded) {
    const std::string state = std::string(argv[2]) + ".state";
    if (!hair.save_state(state.c_str()) || !hair.load_state(state.c_str())) return 3;
    std::remove(state.c_str());
    if (hair.stream(stream4b.data()) != nstream || stream4 != stream4b) return 3;
  }
  f = std::fopen(argv[2], "wb");
  if (!f) return 2;
  std::fwrite(&V, 8, 1, f); std::fwrite(&nelems, 8, 1, f);
  std::fwrite(p4.data(), 4, p4.size(), f); std::fwrite(v4.data(), 4, v4.size(), f); std::fwrite(t4.data(), 4, t4.size(), f);
  std::fwrite(hair.patch_indices().data(), 4, hair.patch_indices().size(), f);
  std::fwrite(&nstream, 8, 1, f); std::fwrite(stream4.data(), 4, 4 * static_cast<size_t>(nstream), f);
  std::fclose(f);
  std::printf("hair_adaptor_main: %lld strands x %d control points, %d frames, tangent plane at byte %llu, %lld patch elements\n",
              (long long)S, N, nframes, (unsigned long long)hair.tangent_plane_offset(), (long long)nelems);
  hair.deinit();
  return hair.initialized() ? 3 : 0;
}
