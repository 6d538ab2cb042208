// libstdc++'s own std::random_shuffle + glibc rand(), to pin the oracle's / host mirror's restatement
// (feature_align.cc:53,103 use exactly this pair).  Needs -std=c++14 (removed in C++17).
#include <algorithm>
#include <cstdlib>
#include <vector>
extern "C" void shim_std_shuffle(int n, int* v, unsigned seed, int rounds) {
  std::srand(seed);
  std::vector<int> a(v, v + n);
  for (int r = 0; r < rounds; r++) std::random_shuffle(a.begin(), a.end());
  std::copy(a.begin(), a.end(), v);
}
