// tests/registry_check.cpp -- IdRegistry (the host registry of target ids, include/target_estimation_b200/target_manager.hpp) against
// std::map<unsigned, uint8_t>, the container it replaces, on random single and batched operations.  No GPU.
#include <cstdio>
#include <map>
#include <random>
#include <vector>

#include "target_estimation_b200/target_manager.hpp"

using target_estimation_b200::IdRegistry;

static bool same(IdRegistry& r, const std::map<unsigned, uint8_t>& m) {
  if (r.size() != m.size()) return false;
  auto it = m.begin();
  for (auto const& kv : r) {
    if (kv.first != it->first || kv.second != it->second) return false;
    ++it;
  }
  return true;
}

int main() {
  std::mt19937 gen(7);
  IdRegistry r;
  std::map<unsigned, uint8_t> m;
  int bad = 0;
  for (int round = 0; round < 400; ++round) {
    const int op = (int)(gen() % 6);
    if (op == 0) {   // single insert / overwrite
      unsigned id = gen() % 5000; uint8_t t = (uint8_t)(gen() % 4);
      r[id] = t; m[id] = t;
    } else if (op == 1) {   // single erase by key
      unsigned id = gen() % 5000;
      bad += r.erase(id) != m.erase(id);
    } else if (op == 2) {   // find + erase by iterator
      unsigned id = gen() % 5000;
      auto a = r.find(id); auto b = m.find(id);
      bad += (a == r.end()) != (b == m.end());
      if (a != r.end()) { bad += a->second != b->second; r.erase(a); m.erase(b); }
    } else if (op == 3) {   // ascending batch insert (existing ids keep their type)
      std::map<unsigned, uint8_t> add;
      for (int k = 0; k < 300; ++k) add[gen() % 6000] = 3;
      std::vector<uint32_t> ids;
      for (auto const& kv : add) ids.push_back(kv.first);
      r.insertSorted(ids.data(), ids.size(), 3);
      m.insert(add.begin(), add.end());
    } else if (op == 4) {   // ascending batch erase, unknown ids included
      std::map<unsigned, uint8_t> del;
      for (int k = 0; k < 300; ++k) del[gen() % 6000] = 0;
      std::vector<uint32_t> ids;
      for (auto const& kv : del) { ids.push_back(kv.first); m.erase(kv.first); }
      r.eraseSorted(ids.data(), ids.size());
    } else {   // range insert of pairs + count
      std::map<unsigned, uint8_t> add;
      for (int k = 0; k < 50; ++k) add[10000 + round * 64 + k] = 1;   // beyond the end: the append path
      r.insert(add.begin(), add.end());
      m.insert(add.begin(), add.end());
      unsigned id = gen() % 6000;
      bad += r.count(id) != m.count(id);
    }
    if (!same(r, m)) { ++bad; std::printf("mismatch after round %d (op %d)\n", round, op); break; }
  }
  std::printf("%s: %zu ids, %d problems\n", bad ? "FAILED" : "ok", r.size(), bad);
  return bad ? 1 : 0;
}
