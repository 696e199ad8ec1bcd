// Test driver for the IqDataProcessor drop-in's public Fs/4 helpers and for what acceptIqData
// leaves in the caller's buffer: argv[1] = "up" | "down" | "accept"; stdin -> stdout.
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "IqDataProcessor.h"

int32_t radio_adjustableReceiveGainInDb = 0;
void nprintf(FILE *s, const char *formatPtr, ...)
{
  va_list ap;
  va_start(ap, formatPtr);
  vfprintf(s, formatPtr, ap);
  va_end(ap);
}

int main(int argc, char **argv)
{
  std::vector<unsigned char> buf;
  unsigned char tmp[4096];
  size_t n;
  while ((n = fread(tmp, 1, sizeof tmp, stdin)) > 0) buf.insert(buf.end(), tmp, tmp + n);
  static char host[] = "127.0.0.1";
  IqDataProcessor p(host, 8001);
  if (argc > 1 && !strcmp(argv[1], "up")) p.upconvertByFsOver4((int8_t *)buf.data(), (uint32_t)buf.size());
  else if (argc > 1 && !strcmp(argv[1], "down")) p.downconvertByFsOver4((int8_t *)buf.data(), (uint32_t)buf.size());
  else p.acceptIqData(0, buf.data(), buf.size());  // mode None: no demodulator, the buffer is still converted
  fwrite(buf.data(), 1, buf.size(), stdout);
  return 0;
}
