// test_host_mirror.cpp -- the reference's own tests, restated against the C++ host mirror (include/pmt.hpp) over libpmt.
//
// Mirrors /root/reference/src/simple_merkle_tree/simple_merkle_tree.rs:117-310 (test_build_merkle_tree_4_leaves,
// test_build_merkle_tree_16_leaves, test_merkle_proof_small_tree, test_verify_small_merkle_proof,
// test_verify_merkle_proof_16) and /root/reference/src/mmr/merkle_mountain_ranges.rs:278-374 (test_heights_bitmap,
// test_get_mmr_index, test_mmr_add_leaf, test_get_proof), with the known answers those tests hold, and then checks the
// mirror against the CPU oracle (oracle/libpmt_oracle.so: test infrastructure, the checker only) on seeded inputs.
// Needs a CUDA device: exits with status 3 and a message when libpmt finds none.  Run by tests/test_cpp_host.py.
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

#include "../../include/pmt.hpp"
#include "../../oracle/pmt_oracle.h"

using namespace pmt;
namespace smt = pmt::simple_merkle_tree;

static int g_checks = 0, g_failed = 0;
#define CHECK(cond)                                                                      \
  do {                                                                                   \
    g_checks++;                                                                          \
    if (!(cond)) { g_failed++; std::printf("  FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); } \
  } while (0)

static HashOut H(uint64_t a, uint64_t b, uint64_t c, uint64_t d) { HashOut h; h.elements = {a, b, c, d}; return h; }
static uint64_t splitmix(uint64_t& s) {
  uint64_t z = (s += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
static std::vector<F> random_felts(size_t n, uint64_t seed, bool canonical = true) {
  std::vector<F> v(n);
  for (auto& x : v) { x = splitmix(seed); if (canonical && x >= 0xFFFFFFFF00000001ull) x -= 0xFFFFFFFF00000001ull; }
  return v;
}
template <class Fn> static bool throws(Fn&& f, int code) {
  try { f(); } catch (const Error& e) { return e.code == code; }
  return false;
}

static const std::vector<F> LEAVES4 = {2890852870ull, 156728478ull, 2876514289ull, 984286162ull};
// simple_merkle_tree.rs:152-167 (from_noncanonical_u128 of values below 2^64)
static const std::vector<F> LEAVES16 = {
    14786323743454721611ull, 976503040092093812ull,   4644130751253292674ull,  6522877527545910706ull,
    11021172818651636092ull, 12048403458499719587ull, 11457874926809001558ull, 14982007443548219923ull,
    4546369223935415035ull,  7205140577604465038ull,  4644130751253292674ull,  4208177174652750506ull,
    16147116534354400672ull, 18147003476480002882ull, 14133393155459789216ull, 9890944065319669426ull};

// simple_merkle_tree.rs:117-143
static void test_build_merkle_tree_4_leaves(const Engine& e) {
  const smt::MerkleTree tree = smt::MerkleTree::build(e, LEAVES4);
  CHECK(tree.count_levels == 2);
  CHECK(tree.tree.size() == 2 && tree.tree[0].size() == 4 && tree.tree[1].size() == 2);
  CHECK(tree.tree[0][1] == H(156728478ull, 0, 0, 0));
  CHECK(tree.tree[1][0] == H(6678006133445961348ull, 15827935749738443865ull, 6295652393730592048ull, 1546515167911236130ull));  // :138
  CHECK(tree.tree[1][1] == H(6698018865469624861ull, 12486244005715193285ull, 11330639022572315007ull, 6059804404595156248ull));
  CHECK(tree.root == H(13451271846715771774ull, 4069913004933160254ull, 14528216580130305557ull, 9716424959297545638ull));      // :140
}

// :145-193
static void test_build_merkle_tree_16_leaves(const Engine& e) {
  const smt::MerkleTree tree = smt::MerkleTree::build(e, LEAVES16);
  CHECK(tree.count_levels == 4);
  CHECK(tree.root == H(2659148958598424285ull, 16496267010313658247ull, 12216516055477211974ull, 15749220035779350537ull));    // :190
}

// :195-213
static void test_merkle_proof_small_tree(const Engine& e) {
  const smt::MerkleTree tree = smt::MerkleTree::build(e, LEAVES4);
  const auto res_leaf_0 = tree.get_merkle_proof(0);
  CHECK(res_leaf_0.size() == 2);
  CHECK(res_leaf_0[0] == H(156728478ull, 0, 0, 0));                                                                             // :210
  CHECK(res_leaf_0[1] == H(6698018865469624861ull, 12486244005715193285ull, 11330639022572315007ull, 6059804404595156248ull)); // :211
}

// :215-235
static void test_verify_small_merkle_proof(const Engine& e) {
  const smt::MerkleTree tree = smt::MerkleTree::build(e, LEAVES4);
  CHECK(smt::verify_merkle_proof(e, LEAVES4[0], 0, tree.root, tree.get_merkle_proof(0)));
  CHECK(smt::verify_merkle_proof(e, LEAVES4[3], 3, tree.root, tree.get_merkle_proof(3)));
}

// :237-310, plus the negative cases a verifier must reject
static void test_verify_merkle_proof_16(const Engine& e) {
  const smt::MerkleTree tree = smt::MerkleTree::build(e, LEAVES16);
  for (size_t i = 0; i < 16; i++) CHECK(smt::verify_merkle_proof(e, LEAVES16[i], i, tree.root, tree.get_merkle_proof(i)));
  CHECK(!smt::verify_merkle_proof(e, LEAVES16[5], 4, tree.root, tree.get_merkle_proof(5)));   // wrong index
  CHECK(!smt::verify_merkle_proof(e, LEAVES16[5] + 1, 5, tree.root, tree.get_merkle_proof(5)));  // wrong leaf
  auto bad = tree.get_merkle_proof(5);
  bad[2].elements[3] ^= 1;
  CHECK(!smt::verify_merkle_proof(e, LEAVES16[5], 5, tree.root, bad));
  const auto between = tree.get_in_between_hashes(5);
  CHECK(between.size() == 4 && between[0] == tree.tree[1][2] && between[1] == tree.tree[2][1] && between[2] == tree.tree[3][0] &&
        between[3] == tree.root);
  // the reference's panics
  CHECK(throws([&] { tree.get_merkle_proof(16); }, PMT_E_RANGE));                                                  // :56
  CHECK(throws([&] { smt::MerkleTree::build(e, std::vector<F>(12, 1)); }, PMT_E_NOT_POW2));                        // :30
  CHECK(throws([&] { smt::MerkleTree::build(e, std::vector<F>(1, 1)); }, PMT_E_INVALID_ARG));                      // :38
}

// merkle_mountain_ranges.rs:278-303
static void test_heights_bitmap() {
  // (mmr_size, bitmap) pairs of the reference's table; the remainder is 0 for each of them
  const std::pair<size_t, uint64_t> table[] = {{1, 1},   {3, 2},   {4, 3},   {7, 4},   {10, 6},  {15, 8},  {22, 12}, {25, 14},
                                               {26, 15}, {31, 16}, {32, 17}, {34, 18}, {35, 19}, {38, 20}, {41, 22}, {42, 23}};
  for (const auto& row : table) {
    CHECK(mmr::get_heights_bitmap_for_mmr_size(row.first).first == row.second);
    CHECK(mmr::get_heights_bitmap_for_mmr_size(row.first).second == 0);
  }
  for (size_t s = 0; s < 5000; s++) {
    uint64_t peaks; size_t rem;
    pmt_oracle_mmr_heights_bitmap(s, &peaks, &rem);
    CHECK(mmr::get_heights_bitmap_for_mmr_size(s) == std::make_pair(peaks, rem));
  }
}

// :305-328
static void test_get_mmr_index() {
  const std::pair<size_t, size_t> table[] = {{0, 0},  {1, 1},  {2, 3},   {3, 4},   {4, 7},   {5, 8},   {6, 10},  {7, 11},
                                             {8, 15}, {9, 16}, {10, 18}, {11, 19}, {12, 22}, {13, 23}, {14, 25}, {15, 26}};
  for (const auto& row : table) CHECK(mmr::get_mmr_index(row.first) == row.second);
  for (size_t i = 0; i < 5000; i++) CHECK(mmr::get_mmr_index(i) == pmt_oracle_mmr_index(i));
  CHECK(throws([] { mmr::get_mmr_index(size_t(1) << 30); }, PMT_E_RANGE));                                         // :264
}

static std::vector<HashOut> oracle_mmr(const std::vector<F>& leaves) {
  std::vector<HashOut> el(2 * leaves.size() + 1);
  size_t len = 0;
  for (F x : leaves) pmt_oracle_mmr_add_leaf(reinterpret_cast<uint64_t*>(el.data()), &len, x);
  el.resize(len);
  return el;
}

// :330-340 -- 100 leaves one at a time, here also compared with the oracle's add_leaf loop
static void test_mmr_add_leaf(const Engine& e) {
  const auto leaves = random_felts(100, 7);
  mmr::MMR m = mmr::MMR::new_();
  for (F x : leaves) m.add_leaf(e, x);
  CHECK(m.elements.size() == 197);   // 2 * 100 - popcount(100)
  CHECK(m.elements == oracle_mmr(leaves));
  mmr::MMR batch = mmr::MMR::new_();
  batch.extend(e, leaves);
  CHECK(batch.elements == m.elements);
}

// :342-374 -- 16 leaves, proof of mmr index 7 = normal index 4; then every leaf of ragged sizes against the oracle
static void test_get_proof(const Engine& e) {
  const auto leaves = random_felts(16, 11);
  mmr::MMR m = mmr::MMR::new_();
  for (F x : leaves) m.add_leaf(e, x);
  const mmr::MMR_proof proof = m.get_proof(7);
  const HashOut root = m.bagging_the_peaks(e);
  CHECK(proof.mmr_size == 31 && proof.merkle_proof.size() == 4 && proof.peaks.size() == 1);
  CHECK(proof.verify(e, leaves[4], root));
  CHECK(!proof.verify(e, leaves[4], H(1, 2, 3, 4)));                                  // wrong root: false
  CHECK(throws([&] { proof.verify(e, leaves[5], root); }, PMT_E_INVALID_ARG));        // assert! at :245
  CHECK(throws([&] { m.get_proof(2); }, PMT_E_INVALID_ARG));                          // position 2 is an inner node

  for (size_t n : {1u, 2u, 3u, 7u, 13u, 64u, 100u, 1000u}) {
    const auto lv = random_felts(n, 100 + n, /*canonical=*/false);
    mmr::MMR r = mmr::MMR::new_();
    r.extend(e, std::vector<F>(lv.begin(), lv.begin() + n / 2));   // two appends: resume from an existing MMR
    r.extend(e, std::vector<F>(lv.begin() + n / 2, lv.end()));
    const auto el = oracle_mmr(lv);
    CHECK(r.elements == el);
    const uint64_t* elw = reinterpret_cast<const uint64_t*>(el.data());
    std::vector<HashOut> opeaks(64);
    opeaks.resize(pmt_oracle_mmr_peaks(elw, el.size(), reinterpret_cast<uint64_t*>(opeaks.data())));
    CHECK(r.get_peaks() == opeaks);
    HashOut oroot;
    pmt_oracle_mmr_bag(elw, el.size(), oroot.elements.data());
    const HashOut bag = r.bagging_the_peaks(e);
    CHECK(bag == oroot);
    for (size_t i = 0; i < n; i += (n > 100 ? 37 : 1)) {
      const mmr::MMR_proof p = r.get_proof_normal_index(i);
      std::vector<HashOut> osib(64);
      std::vector<uint8_t> oleft(64);
      const size_t len = pmt_oracle_mmr_subtree_proof(elw, el.size(), mmr::get_mmr_index(i), reinterpret_cast<uint64_t*>(osib.data()), oleft.data());
      bool same = p.merkle_proof.size() == len && p.mmr_size == el.size();
      for (size_t j = 0; same && j < len; j++) same = p.merkle_proof[j].first == osib[j] && p.merkle_proof[j].second == (oleft[j] != 0);
      CHECK(same);
      CHECK(p.verify(e, lv[i], bag));
    }
  }
}

// [UPSTREAM] MerkleTree::new / prove / verify_merkle_proof_to_cap against the oracle, incl. wide leaves and cap = log2 n
static void test_plonky2_merkle_tree(const Engine& e) {
  struct Shape { size_t n, w, cap; };
  for (const Shape s : {Shape{2, 1, 0}, Shape{16, 4, 0}, Shape{16, 4, 4}, Shape{64, 9, 2}, Shape{256, 135, 4}, Shape{1024, 4, 1}}) {
    const auto flat = random_felts(s.n * s.w, 1000 + s.n + s.w, false);
    std::vector<std::vector<F>> leaves(s.n);
    for (size_t i = 0; i < s.n; i++) leaves[i].assign(flat.begin() + i * s.w, flat.begin() + (i + 1) * s.w);
    const plonky2::MerkleTree t = plonky2::MerkleTree::new_(e, leaves, s.cap);
    std::vector<HashOut> dig(2 * (s.n - (size_t(1) << s.cap)) + 1), cap(size_t(1) << s.cap);
    CHECK(pmt_oracle_merkle_tree_new(flat.data(), s.n, s.w, (unsigned)s.cap, reinterpret_cast<uint64_t*>(dig.data()),
                                     reinterpret_cast<uint64_t*>(cap.data()), 1, 0) == 0);
    dig.resize(dig.size() - 1);
    CHECK(t.digests == dig);
    CHECK(t.cap.hashes == cap);
    for (size_t i = 0; i < s.n; i += (s.n > 64 ? 17 : 1)) {
      const plonky2::MerkleProof p = t.prove(i);
      std::vector<HashOut> osib(p.siblings.size() + 1);
      CHECK(pmt_oracle_merkle_prove(reinterpret_cast<const uint64_t*>(dig.data()), s.n, (unsigned)s.cap, i, reinterpret_cast<uint64_t*>(osib.data())) == 0);
      osib.resize(p.siblings.size());
      CHECK(p.siblings == osib);
      CHECK(plonky2::verify_merkle_proof_to_cap(e, leaves[i], i, t.cap, p));
      CHECK(!plonky2::verify_merkle_proof_to_cap(e, leaves[i], i ^ 1, t.cap, p) || s.n == (size_t(1) << s.cap));
    }
  }
  CHECK(throws([&] { plonky2::MerkleTree::new_(e, std::vector<std::vector<F>>(8, std::vector<F>(4, 1)), 4); }, PMT_E_RANGE));
  CHECK(throws([&] { plonky2::MerkleTree::new_(e, std::vector<std::vector<F>>(6, std::vector<F>(4, 1)), 0); }, PMT_E_NOT_POW2));
}

// MerkleTree::new over several contexts in one process (here all on device 0): equal to the single-context tree
static void test_plonky2_merkle_tree_multi(const Engine& e) {
  Engine e1, e2, e3;
  const std::vector<const Engine*> four{&e, &e1, &e2, &e3}, two{&e2, &e3};
  struct Shape { size_t n, w, cap; };
  for (const Shape s : {Shape{4, 4, 0}, Shape{16, 1, 0}, Shape{16, 4, 2}, Shape{64, 9, 1}, Shape{64, 135, 4}, Shape{4096, 4, 0}}) {
    const auto flat = random_felts(s.n * s.w, 2000 + s.n + s.w, true);
    std::vector<std::vector<F>> leaves(s.n);
    for (size_t i = 0; i < s.n; i++) leaves[i].assign(flat.begin() + i * s.w, flat.begin() + (i + 1) * s.w);
    const plonky2::MerkleTree one = plonky2::MerkleTree::new_(e, leaves, s.cap);
    for (const auto* pool : {&four, &two}) {
      const plonky2::MerkleTree t = plonky2::MerkleTree::new_multi(*pool, leaves, s.cap);
      CHECK(t.digests == one.digests);
      CHECK(t.cap.hashes == one.cap.hashes);
    }
  }
  // MMR batch append over several contexts: equal to the single-context append, also onto a non-empty, unaligned MMR
  for (const size_t n0 : {size_t(0), size_t(777), size_t(8192)}) {
    const auto lv = random_felts(n0 + 3 * 8192 + 11, 3000 + n0, false);
    const std::vector<F> head(lv.begin(), lv.begin() + n0), rest(lv.begin() + n0, lv.end());
    mmr::MMR a = mmr::MMR::new_(), b = mmr::MMR::new_();
    a.extend(e, head); a.extend(e, rest);
    b.extend(e, head); b.extend_multi(two, rest);
    CHECK(a.elements == b.elements);
    CHECK(a.bagging_the_peaks(e) == b.bagging_the_peaks(e));
  }
  CHECK(throws([&] { plonky2::MerkleTree::new_multi({&e, &e}, std::vector<std::vector<F>>(8, std::vector<F>(4, 1)), 0); }, PMT_E_INVALID_ARG));
  CHECK(throws([&] { plonky2::MerkleTree::new_multi({&e, &e1, &e2}, std::vector<std::vector<F>>(8, std::vector<F>(4, 1)), 0); }, PMT_E_NOT_POW2));
  CHECK(throws([&] { plonky2::MerkleTree::new_multi(four, std::vector<std::vector<F>>(2, std::vector<F>(4, 1)), 0); }, PMT_E_RANGE));
}

// Hasher: the asserted known answers of simple_merkle_tree.rs:210-211 and the oracle on random inputs
static void test_hasher(const Engine& e) {
  CHECK(e.hash_or_noop({156728478ull}) == H(156728478ull, 0, 0, 0));
  CHECK(e.two_to_one(H(2876514289ull, 0, 0, 0), H(984286162ull, 0, 0, 0)) ==
        H(6698018865469624861ull, 12486244005715193285ull, 11330639022572315007ull, 6059804404595156248ull));
  for (size_t w : {1u, 4u, 5u, 8u, 9u, 12u, 96u, 135u}) {
    const auto in = random_felts(w, 50 + w, false);
    HashOut o;
    pmt_oracle_hash_or_noop(in.data(), w, o.elements.data());
    CHECK(e.hash_or_noop(in) == o);
    pmt_oracle_hash_no_pad(in.data(), w, o.elements.data());
    CHECK(e.hash_no_pad(in) == o);
  }
}

int main() {
  std::printf("test_heights_bitmap\n");  test_heights_bitmap();     // pure index math: runs without a device
  std::printf("test_get_mmr_index\n");   test_get_mmr_index();
  if (g_failed) return 1;
  Engine* engine = nullptr;
  try { engine = new Engine(0); } catch (const Error& err) {
    std::printf("no engine: %s (index-math tests passed: %d checks)\n", err.what(), g_checks);
    return 3;
  }
  const Engine& e = *engine;
  const std::pair<const char*, std::function<void()>> tests[] = {
      {"test_hasher", [&] { test_hasher(e); }},
      {"test_build_merkle_tree_4_leaves", [&] { test_build_merkle_tree_4_leaves(e); }},
      {"test_build_merkle_tree_16_leaves", [&] { test_build_merkle_tree_16_leaves(e); }},
      {"test_merkle_proof_small_tree", [&] { test_merkle_proof_small_tree(e); }},
      {"test_verify_small_merkle_proof", [&] { test_verify_small_merkle_proof(e); }},
      {"test_verify_merkle_proof_16", [&] { test_verify_merkle_proof_16(e); }},
      {"test_mmr_add_leaf", [&] { test_mmr_add_leaf(e); }},
      {"test_get_proof", [&] { test_get_proof(e); }},
      {"test_plonky2_merkle_tree", [&] { test_plonky2_merkle_tree(e); }},
      {"test_plonky2_merkle_tree_multi", [&] { test_plonky2_merkle_tree_multi(e); }},
  };
  for (const auto& t : tests) {
    std::printf("%s\n", t.first);
    try { t.second(); } catch (const std::exception& ex) { g_failed++; std::printf("  EXCEPTION: %s\n", ex.what()); }
  }
  std::printf("%d checks, %d failed, %" PRIu64 " kernel launches\n", g_checks, g_failed, e.kernel_launches());
  delete engine;
  return g_failed ? 1 : 0;
}
