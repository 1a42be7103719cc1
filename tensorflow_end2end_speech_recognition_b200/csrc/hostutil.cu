// Host-side helpers behind the C ABI (no device code).
//   b2_crc32c: CRC-32C (Castagnoli, reflected polynomial 0x82F63B78), the checksum of TensorFlow's checkpoint bundles
//   (every table block of the .index file and every tensor of the .data shard carries one;
//   utils/io/tf_checkpoint.py).  Slicing-by-8; a pure-Python loop would need minutes for the 110 MB of a config-2 model.
#include "common.cuh"
#include <string.h>

namespace {
uint32_t g_tab[8][256];
bool g_tab_ready = false;
void crc_init() {
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
    g_tab[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i)
    for (int t = 1; t < 8; ++t) g_tab[t][i] = (g_tab[t - 1][i] >> 8) ^ g_tab[0][g_tab[t - 1][i] & 0xff];
  g_tab_ready = true;
}
}  // namespace

// crc: the value returned for the preceding bytes (0 for the first call); returns the CRC-32C of the concatenation
extern "C" uint32_t b2_crc32c(uint32_t crc, const void* data, size_t n) {
  if (!g_tab_ready) crc_init();
  const unsigned char* p = (const unsigned char*)data;
  uint32_t c = ~crc;
  while (n && ((uintptr_t)p & 7)) { c = g_tab[0][(c ^ *p++) & 0xff] ^ (c >> 8); --n; }
  while (n >= 8) {
    uint64_t v;
    memcpy(&v, p, 8);
    v ^= c;
    c = g_tab[7][v & 0xff] ^ g_tab[6][(v >> 8) & 0xff] ^ g_tab[5][(v >> 16) & 0xff] ^ g_tab[4][(v >> 24) & 0xff] ^
        g_tab[3][(v >> 32) & 0xff] ^ g_tab[2][(v >> 40) & 0xff] ^ g_tab[1][(v >> 48) & 0xff] ^ g_tab[0][v >> 56];
    p += 8; n -= 8;
  }
  while (n--) c = g_tab[0][(c ^ *p++) & 0xff] ^ (c >> 8);
  return ~c;
}
