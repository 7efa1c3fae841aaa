// Host scalar inversion (csrc/host/sc_host.hpp): the binary inverse the prover uses for its challenges against the
// exponentiation x^(l-2) (the reference's Scalar::invert) and against x * x^-1 = 1.  Prints the mismatch count.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include "../../mpc_bulletproof_b200/csrc/host/sc_host.hpp"
using namespace bpg_host;
int main() {
  std::mt19937_64 g(20240607);
  int bad = 0;
  for (int i = 0; i < 60000; i++) {
    uint8_t b[64];
    for (int j = 0; j < 64; j += 8) {
      uint64_t x = g();
      memcpy(b + j, &x, 8);
    }
    Scalar x = Scalar::from_wide(b);
    if (i < 64) x = Scalar::from_u64((uint64_t)i);               // 0, 1, 2, ... small values
    else if (i < 128) x = -Scalar::from_u64((uint64_t)(i - 63));  // l - 1, l - 2, ...
    else if (i < 192) {                                           // powers of two (long runs of halvings)
      x = Scalar::one();
      for (int k = 0; k < 4 * (i - 128); k++) x = x + x;
    }
    Scalar a = x.invert();
    if (i < 4000 || i % 8 == 0)
      if (!(a == x.invert_fermat())) bad++;
    if (!x.is_zero() && !((a * x) == Scalar::one())) bad++;
    if (x.is_zero() && !a.is_zero()) bad++;
  }
  printf("mismatches %d\n", bad);
  return bad != 0;
}
