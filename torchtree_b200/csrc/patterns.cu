// Site-pattern compression (host code): unique alignment columns, sorted the way
// the reference sorts them, with their multiplicities.
//
// Replaces the Python `Counter(zip(*sequences))` + `sorted(...)` of
// torchtree/evolution/site_pattern.py:69-97 (`compress`), which walks every
// character of the alignment in the interpreter (minutes and GBs of tuples for a
// 1000 x 100k alignment).  SURVEY 8(f) row f3.
//
// A site is `group` consecutive characters per sequence (1 for nucleotides and
// amino acids, 3 for codons, site_pattern.py:79-82).  The reference orders
// patterns as Python tuples of per-taxon strings (or tuples of characters), i.e.
// lexicographically over the taxa and, within a taxon, over the characters of
// the group -- exactly memcmp over the site's key laid out [taxon][char].
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "engine.cuh"

namespace {

// Run body(begin, end) over [0, count) split across the host's hardware threads.
template <class F>
void parallel_blocks(int64_t count, int64_t grain, F body) {
  unsigned hw = std::thread::hardware_concurrency();
  int64_t workers = std::max<int64_t>(1, std::min<int64_t>(hw ? hw : 1, (count + grain - 1) / grain));
  if (workers == 1) {
    body((int64_t)0, count);
    return;
  }
  const int64_t per = ((count + workers - 1) / workers + grain - 1) / grain * grain;
  std::vector<std::thread> pool;
  for (int64_t w = 0; w < workers; ++w) {
    const int64_t b = w * per, e = std::min(count, b + per);
    if (b >= e) break;
    pool.emplace_back([=] { body(b, e); });
  }
  for (auto& th : pool) th.join();
}

}  // namespace

extern "C" int ttb2_compress_patterns(const uint8_t* sequences, int32_t taxa, int64_t length,
                                      int32_t group, uint8_t* patterns, double* weights,
                                      int64_t* pattern_count) {
  using ttb2::set_error;
  if (!sequences || !patterns || !weights || !pattern_count) {
    set_error("ttb2_compress_patterns: null argument");
    return TTB2_E_INVALID;
  }
  if (taxa < 1 || group < 1 || length < group || length % group != 0) {
    set_error("ttb2_compress_patterns: need taxa >= 1, group >= 1 and length a multiple of group");
    return TTB2_E_INVALID;
  }
  const bool timing = std::getenv("TTB2_PATTERNS_TIMING") != nullptr;
  auto t0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    auto t1 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[ttb2 patterns] %s %.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
    t0 = t1;
  };
  const int64_t sites = length / group;
  const size_t keyLen = (size_t)taxa * group;
  // site-major keys [site][taxon][char]
  std::unique_ptr<uint8_t[]> keys(new uint8_t[(size_t)sites * keyLen]);
  // blocked transpose so that both sides stay in cache, sites split over threads
  constexpr int64_t TILE = 64;
  parallel_blocks(sites, TILE, [&](int64_t jb, int64_t je) {
    for (int64_t j0 = jb; j0 < je; j0 += TILE) {
      const int64_t j1 = std::min(je, j0 + TILE);
      for (int32_t t0 = 0; t0 < taxa; t0 += (int32_t)TILE) {
        const int32_t t1 = (int32_t)std::min<int64_t>(taxa, t0 + TILE);
        for (int32_t t = t0; t < t1; ++t) {
          const uint8_t* src = sequences + (size_t)t * length;
          uint8_t* dst = keys.get() + (size_t)t * group;
          if (group == 1) {
            for (int64_t j = j0; j < j1; ++j) dst[(size_t)j * keyLen] = src[j];
          } else {
            for (int64_t j = j0; j < j1; ++j)
              for (int32_t g = 0; g < group; ++g) dst[(size_t)j * keyLen + g] = src[j * group + g];
          }
        }
      }
    }
  });
  lap("transpose in");
  std::vector<int64_t> order(sites);
  std::iota(order.begin(), order.end(), (int64_t)0);
  const uint8_t* base = keys.get();
  std::sort(order.begin(), order.end(), [base, keyLen](int64_t a, int64_t b) {
    const int c = std::memcmp(base + (size_t)a * keyLen, base + (size_t)b * keyLen, keyLen);
    return c < 0 || (c == 0 && a < b);
  });
  lap("sort");
  // unique + counts; patterns are written [taxon][pattern][char] so that one
  // taxon's row is contiguous (the engine's tip layout)
  int64_t n = 0;
  for (int64_t r = 0; r < sites; ++r) {
    const uint8_t* k = base + (size_t)order[r] * keyLen;
    if (r > 0 && std::memcmp(k, base + (size_t)order[r - 1] * keyLen, keyLen) == 0) {
      weights[n - 1] += 1.0;
      continue;
    }
    weights[n] = 1.0;
    ++n;
  }
  // second pass to place the patterns with the final stride n
  std::vector<int64_t> first;
  first.reserve(n);
  for (int64_t r = 0; r < sites; ++r) {
    const uint8_t* k = base + (size_t)order[r] * keyLen;
    if (r > 0 && std::memcmp(k, base + (size_t)order[r - 1] * keyLen, keyLen) == 0) continue;
    first.push_back(order[r]);
  }
  parallel_blocks(n, TILE, [&](int64_t pb, int64_t pe) {
    for (int64_t p0 = pb; p0 < pe; p0 += TILE) {
      const int64_t p1 = std::min(pe, p0 + TILE);
      for (int32_t t0 = 0; t0 < taxa; t0 += (int32_t)TILE) {
        const int32_t t1 = (int32_t)std::min<int64_t>(taxa, t0 + TILE);
        for (int32_t t = t0; t < t1; ++t) {
          uint8_t* dst = patterns + (size_t)t * n * group;
          for (int64_t p = p0; p < p1; ++p) {
            const uint8_t* k = base + (size_t)first[p] * keyLen + (size_t)t * group;
            for (int32_t g = 0; g < group; ++g) dst[p * group + g] = k[g];
          }
        }
      }
    }
  });
  lap("unique + transpose out");
  *pattern_count = n;
  return TTB2_OK;
}
