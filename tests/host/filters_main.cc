// Test driver for the filter-class drop-ins (rtlsdrdiags_b200/host/B200Filters.h):
//   filters_main <dec|int|fir><32|16> <factor> <taps file (float32)> <block|sample>   stdin -> stdout
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "B200Filters.h"

template <class T>
static std::vector<T> slurp(FILE *f)
{
  std::vector<T> v;
  T buf[1024];
  size_t n;
  while ((n = fread(buf, sizeof(T), 1024, f)) > 0) v.insert(v.end(), buf, buf + n);
  return v;
}

template <class T, class DEC, class INT, class FIR>
static int run(const char *what, int factor, std::vector<float> &h, bool block)
{
  std::vector<T> x = slurp<T>(stdin), y;
  const uint32_t n = (uint32_t)x.size();
  if (!strncmp(what, "dec", 3))
  {
    DEC d((int)h.size(), h.data(), factor);
    if (d.lastStatus() != SDR_OK) return 2;
    y.resize(n / factor + 2);
    uint32_t c = 0;
    if (block)
    {  // two uneven pieces: the samples waiting between calls must carry over
      c = d.decimate(x.data(), n / 3, y.data());
      c += d.decimate(x.data() + n / 3, n - n / 3, y.data() + c);
    }
    else
      for (uint32_t i = 0; i < n; i++) { T o; if (d.decimate(x[i], &o)) y[c++] = o; }
    y.resize(c);
  }
  else if (!strncmp(what, "int", 3))
  {
    INT p((int)h.size(), h.data(), factor);
    if (p.lastStatus() != SDR_OK) return 2;
    y.resize((size_t)n * factor);
    if (block) { p.interpolate(x.data(), n / 2, y.data()); p.interpolate(x.data() + n / 2, n - n / 2, y.data() + (size_t)(n / 2) * factor); }
    else for (uint32_t i = 0; i < n; i++) p.interpolate(x[i], &y[(size_t)i * factor]);
  }
  else
  {
    FIR f((int)h.size(), h.data());
    if (f.lastStatus() != SDR_OK) return 2;
    y.resize(n);
    if (block) f.filterData(x.data(), n, y.data());
    else for (uint32_t i = 0; i < n; i++) y[i] = f.filterData(x[i]);
  }
  fwrite(y.data(), sizeof(T), y.size(), stdout);
  return 0;
}

int main(int argc, char **argv)
{
  if (argc < 5) return 1;
  FILE *tf = fopen(argv[3], "rb");
  if (!tf) return 1;
  std::vector<float> h = slurp<float>(tf);
  fclose(tf);
  const bool block = !strcmp(argv[4], "block");
  const int factor = atoi(argv[2]);
  if (strstr(argv[1], "16")) return run<int16_t, Decimator_int16, Interpolator_int16, FirFilter_int16>(argv[1], factor, h, block);
  return run<float, Decimator, Interpolator, FirFilter>(argv[1], factor, h, block);
}
