// Stress driver for csrc/runtime/data_loader.cc, built with -fsanitize=thread / address by bench/sanitize_native.sh:
// several loaders with different geometries run at once over shared token files, consumers of different speeds, a loader
// stopped while its workers are mid-batch, a source whose files are too short (exception on a worker thread -> re-raised by Acquire).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "runtime/data_loader.h"

using namespace tepdist;

static std::string WriteFile(const std::string& path, int n, int base) {
  std::vector<uint16_t> v((size_t)n);
  for (int i = 0; i < n; ++i) v[i] = (uint16_t)((base + i) & 0xFFFF);
  std::ofstream f(path, std::ios::binary);
  f.write((const char*)v.data(), (std::streamsize)(v.size() * 2));
  return path;
}

static int RunLoader(std::shared_ptr<TokenSource> src, int batch, int ctx, int rank, int world, int slots, int threads, int steps,
                     int consumer_sleep_us, bool stop_early) {
  BatchLoader ld(src, batch, ctx, rank, world, /*seed=*/7, threads);
  std::vector<std::vector<int32_t>> tok((size_t)slots, std::vector<int32_t>((size_t)batch * ctx)), lab = tok;
  std::vector<uintptr_t> tp, lp;
  for (int s = 0; s < slots; ++s) { tp.push_back((uintptr_t)tok[s].data()); lp.push_back((uintptr_t)lab[s].data()); }
  ld.SetBuffers(tp, lp);
  ld.Start(0);
  int bad = 0, held = -1;
  for (int t = 0; t < steps; ++t) {
    if (held >= 0) ld.Release(held);
    uint64_t step = 0;
    held = ld.Acquire(&step);
    if (step != (uint64_t)t) ++bad;
    // every row is a window of consecutive tokens; labels are the next token
    for (int b = 0; b < batch; ++b)
      for (int i = 0; i + 1 < ctx; ++i) {
        const int32_t a = tok[held][(size_t)b * ctx + i], n = tok[held][(size_t)b * ctx + i + 1];
        if (((a + 1) & 0xFFFF) != n || lab[held][(size_t)b * ctx + i] != n) ++bad;
      }
    if (consumer_sleep_us) std::this_thread::sleep_for(std::chrono::microseconds(consumer_sleep_us));
    if (stop_early && t == steps / 2) break;
  }
  ld.Stop();
  return bad;
}

int main(int argc, char** argv) {
  const std::string dir = argc > 1 ? argv[1] : "/tmp";
  auto src = std::make_shared<TokenSource>();
  src->AddDataset({WriteFile(dir + "/s0.bin", 20000, 0), WriteFile(dir + "/s1.bin", 30000, 1000)}, 2.0, 2);
  src->AddDataset({WriteFile(dir + "/s2.bin", 5000, 7)}, 1.0, 2);
  int bad = 0;
  std::vector<std::thread> jobs;
  std::vector<int> res(6, 0);
  jobs.emplace_back([&] { res[0] = RunLoader(src, 8, 64, 0, 1, 4, 3, 200, 0, false); });
  jobs.emplace_back([&] { res[1] = RunLoader(src, 2, 128, 1, 4, 2, 4, 150, 50, false); });
  jobs.emplace_back([&] { res[2] = RunLoader(src, 4, 32, 3, 4, 3, 1, 300, 0, false); });
  jobs.emplace_back([&] { res[3] = RunLoader(src, 4, 256, 0, 2, 8, 4, 100, 200, true); });    // stopped with batches in flight
  for (auto& j : jobs) j.join();
  for (int r : res) bad += r;
  // worker-thread exception: every file shorter than one window
  auto tiny = std::make_shared<TokenSource>();
  tiny->AddDataset({WriteFile(dir + "/tiny.bin", 10, 0)}, 1.0, 2);
  bool raised = false;
  try {
    RunLoader(tiny, 1, 64, 0, 1, 2, 2, 3, 0, false);
  } catch (const std::runtime_error& e) {
    raised = std::strstr(e.what(), "shorter than one window") != nullptr;
  }
  if (!raised) ++bad;
  std::printf("data_loader_stress: %s (%d violations)\n", bad == 0 ? "OK" : "FAILED", bad);
  return bad == 0 ? 0 : 1;
}
