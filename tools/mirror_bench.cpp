// mirror_bench.cpp -- end-to-end time of the headline workload through the C++ host mirror (include/pmt.hpp), i.e. WITH the
// conversions a Rust shim pays around the C ABI call: pmt::plonky2::MerkleTree::new_ takes Vec<Vec<F>> leaves, flattens them
// (512 MiB host copy), sizes the Vec<HashOut> digests (1 GiB, value-initialised: the first touch of fresh pages) and calls
// pmt_merkle_tree_build on those PAGEABLE vectors.  Prints one JSON line: whole-call times, the bare C ABI call on the same
// (pageable) flat vectors, and the root for the caller's parity check.
// build: g++ -std=c++17 -O2 -Iinclude -o tools/_ab/mirror_bench tools/mirror_bench.cpp -Lplonky2_merkle_trees_b200 -lpmt -Wl,-rpath,$ORIGIN/../../plonky2_merkle_trees_b200
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../include/pmt.hpp"

static uint64_t splitmix(uint64_t idx) {   // bench.py's generator, seed 0x706d745f62323030
  const uint64_t P = 0xFFFFFFFF00000001ull;
  uint64_t z = 0x706D745F62323030ull + (idx + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return z >= P ? z - P : z;
}
static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
  const int lg = argc > 1 ? atoi(argv[1]) : 24, reps = argc > 2 ? atoi(argv[2]) : 3;
  const size_t n = (size_t)1 << lg, w = 4;
  try {
    pmt::Engine gpu(0);
    std::vector<std::vector<pmt::F>> leaves(n, std::vector<pmt::F>(w));
    for (size_t i = 0; i < n; i++)
      for (size_t j = 0; j < w; j++) leaves[i][j] = splitmix(i * w + j);
    std::vector<double> whole, bare;
    uint64_t root[4] = {0, 0, 0, 0};
    for (int r = 0; r < reps + 1; r++) {
      auto copy = leaves;                                   // new_ takes the leaves by value (it keeps them, like upstream)
      const double t0 = now_ms();
      auto tree = pmt::plonky2::MerkleTree::new_(gpu, std::move(copy), 0);
      const double t1 = now_ms();
      if (r) whole.push_back(t1 - t0);                      // r == 0: warm-up (arenas, staging slots, copy threads)
      for (int j = 0; j < 4; j++) root[j] = tree.cap.hashes[0].elements[j];
    }
    std::vector<uint64_t> flat(n * w), dig((2 * n - 2) * 4), cap(4);
    for (size_t i = 0; i < n * w; i++) flat[i] = splitmix(i);
    for (int r = 0; r < reps + 1; r++) {
      const double t0 = now_ms();
      gpu.check(pmt_merkle_tree_build(gpu.ctx(), flat.data(), n, w, 0, dig.data(), cap.data()));
      const double t1 = now_ms();
      if (r) bare.push_back(t1 - t0);
    }
    std::sort(whole.begin(), whole.end());
    std::sort(bare.begin(), bare.end());
    printf("{\"mirror\": \"pmt::plonky2::MerkleTree::new_ (include/pmt.hpp)\", \"log2_leaves\": %d, \"ms_median\": %.3f, \"ms_best\": %.3f, "
           "\"c_abi_on_pageable_vectors_ms_median\": %.3f, \"c_abi_on_pageable_vectors_ms_best\": %.3f, \"root\": [%llu, %llu, %llu, %llu], "
           "\"root_matches_bare_call\": %s}\n",
           lg, whole[whole.size() / 2], whole[0], bare[bare.size() / 2], bare[0], (unsigned long long)root[0], (unsigned long long)root[1],
           (unsigned long long)root[2], (unsigned long long)root[3],
           (cap[0] == root[0] && cap[1] == root[1] && cap[2] == root[2] && cap[3] == root[3]) ? "true" : "false");
  } catch (const std::exception& e) {
    fprintf(stderr, "mirror_bench: %s\n", e.what());
    return 1;
  }
  return 0;
}
