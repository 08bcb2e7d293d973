#include "../../rabbitvar_b200/csrc/host/fmt.hpp"
#include <random>
#include <string>
int main() {
  std::mt19937_64 rng(12345);
  long bad = 0, n = 0;
  auto check = [&](double x) {
    std::string a = std::to_string(x), b = rvhost::f6(x);
    ++n;
    if (a != b) { if (bad < 10) printf("MISMATCH %.17g: %s vs %s\n", x, a.c_str(), b.c_str()); ++bad; }
  };
  double specials[] = {0.0, -0.0, 0.5, 1.5, 2.5, 0.0000005, 0.0000015, 0.0000025, 1e-7, -1e-7, 123456.7890125, 0.1, 0.2, 0.3, 1e11, 999999999999.0,
                       1e12, 1e15, -3.25, 60.0, 37.0, 0.020202020202020204, 4.9999995, 0.9999995, 0.9999994999999999, 1.0/3, 2.0/3, 1e-300, 5e-324,
                       INFINITY, -INFINITY, NAN};
  for (double x : specials) check(x);
  std::uniform_real_distribution<double> u(0, 1);
  for (int i = 0; i < 3000000; ++i) {
    double x = u(rng);
    int k = rng() % 8;
    if (k == 0) x *= 100; else if (k == 1) x *= 1e6; else if (k == 2) x = (double)(rng() % 2000) / (double)(1 + rng() % 300);
    else if (k == 3) x = -x * 50; else if (k == 4) x = (rng() % 2000001) / 2e6 + (rng() % 100);   // ties at the 7th decimal
    else if (k == 5) x = ldexp((double)(rng() >> 11), -(int)(rng() % 80));
    check(x);
  }
  printf("checked %ld values, %ld mismatches\n", n, bad);
  return bad != 0;
}
