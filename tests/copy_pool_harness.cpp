// tests/copy_pool_harness.cpp -- stress of ttv_b200/csrc/copy_pool.h on the CPU: several caller threads share one pool
// (what happens when host-pointer calls for different devices run at the same time), random sizes and offsets, every
// copy verified.  Built twice by tests/test_hostcopy.py: plain and with -fsanitize=thread.
#include "../ttv_b200/csrc/copy_pool.h"

#include <cstdio>
#include <cstdlib>
#include <random>

int main(int argc, char** argv)
{
  const int callers = argc > 1 ? std::atoi(argv[1]) : 3, workers = argc > 2 ? std::atoi(argv[2]) : 5, rounds = argc > 3 ? std::atoi(argv[3]) : 40;
  ttvb::CopyPool pool(workers);
  std::atomic<int> bad{0};
  std::vector<std::thread> th;
  for (int c = 0; c < callers; ++c)
    th.emplace_back([&, c] {
      std::mt19937_64 rng(1234 + c);
      const size_t cap = (size_t)24 << 20;
      std::vector<unsigned char> src(cap + 64), dst(cap + 64);
      for (auto& x : src) x = (unsigned char)rng();
      for (int r = 0; r < rounds; ++r) {
        const size_t bytes = r % 7 == 0 ? (size_t)(rng() % 4096) : (size_t)(rng() % cap);      // tiny ones take the shortcut
        const size_t so = rng() % 64, dofs = rng() % 64;
        std::fill(dst.begin(), dst.end(), (unsigned char)0x5A);
        pool.copy(dst.data() + dofs, src.data() + so, bytes);
        if (std::memcmp(dst.data() + dofs, src.data() + so, bytes) != 0) ++bad;
        for (size_t i = 0; i < dofs; ++i) if (dst[i] != 0x5A) ++bad;
        for (size_t i = dofs + bytes; i < dofs + bytes + 32 && i < dst.size(); ++i) if (dst[i] != 0x5A) ++bad;
      }
    });
  for (auto& t : th) t.join();
  std::printf("copy pool: %d callers x %d rounds on %d workers, %d bad\n", callers, rounds, workers, bad.load());
  return bad.load() ? 1 : 0;
}
