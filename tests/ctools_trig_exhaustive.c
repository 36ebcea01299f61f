/* Exhaustive check: cs_sinf/cs_cosf (host build of slam.net_b200/csrc/cs_math.h) == libm sinf/cosf for all
 * 2^32 float bit patterns (NaN results compared as NaN).  Prints the mismatch counts. */
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <string.h>

#include "../slam.net_b200/csrc/cs_math.h"

#define NT 8
static unsigned long long bad[NT];
static unsigned long long stride = 1;

static void* run(void* a) {
  long t = (long)a;
  unsigned long long lo = (unsigned long long)t << 29, hi = lo + (1ull << 29), b = 0;
  for (unsigned long long i = lo; i < hi; i += stride) {
    uint32_t u = (uint32_t)i;
    float f;
    memcpy(&f, &u, 4);
    float s1 = cs_sinf(f), s2 = sinf(f), c1 = cs_cosf(f), c2 = cosf(f);
    if (!((s1 != s1 && s2 != s2) || cs_f2u(s1) == cs_f2u(s2))) b++;
    if (!((c1 != c1 && c2 != c2) || cs_f2u(c1) == cs_f2u(c2))) b++;
  }
  bad[t] = b;
  return 0;
}

int main(int argc, char** argv) {
  if (argc > 1) stride = strtoull(argv[1], 0, 10);
  pthread_t th[NT];
  for (long t = 0; t < NT; t++) pthread_create(&th[t], 0, run, (void*)t);
  unsigned long long total = 0;
  for (int t = 0; t < NT; t++) { pthread_join(th[t], 0); total += bad[t]; }
  printf("mismatches %llu stride %llu\n", total, stride);
  return total != 0;
}
