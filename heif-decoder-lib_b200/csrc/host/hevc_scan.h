// hevc_scan.h — coefficient scan orders (H.265 §6.5.3-6.5.5), generated once at start-up.
// Reference counterpart: third-party/libde265/libde265/scan.cc.
#pragma once
#include <cstdint>

namespace hc {

struct ScanPos { uint8_t x, y; };

struct ScanTables {
  // [log2BlockSize 0..5][scanIdx 0 diag,1 horizontal,2 vertical] -> positions (blk*blk entries)
  ScanPos* order[6][3];
  ScanPos storage[3 * (1 + 4 + 16 + 64 + 256 + 1024)];
  // inverse[log2][scanIdx][x + (y << log2)] -> scan index, for log2 0..3 (sub-block grids and the 4x4 positions)
  uint8_t inverse[4][3][64];
  ScanTables() {
    ScanPos* p = storage;
    for (int l = 0; l <= 5; l++) {
      int n = 1 << l;
      // diagonal (up-right)
      order[l][0] = p;
      {
        int i = 0, x = 0, y = 0;
        bool stop = false;
        while (!stop) {
          while (y >= 0) {
            if (x < n && y < n) { p[i].x = (uint8_t)x; p[i].y = (uint8_t)y; i++; }
            y--; x++;
          }
          y = x; x = 0;
          if (i >= n * n) stop = true;
        }
        p += n * n;
      }
      order[l][1] = p;
      for (int y = 0, i = 0; y < n; y++)
        for (int x = 0; x < n; x++, i++) { p[i].x = (uint8_t)x; p[i].y = (uint8_t)y; }
      p += n * n;
      order[l][2] = p;
      for (int x = 0, i = 0; x < n; x++)
        for (int y = 0; y < n; y++, i++) { p[i].x = (uint8_t)x; p[i].y = (uint8_t)y; }
      p += n * n;
    }
    for (int l = 0; l <= 3; l++)
      for (int s = 0; s < 3; s++)
        for (int i = 0; i < (1 << (2 * l)); i++) inverse[l][s][order[l][s][i].x + (order[l][s][i].y << l)] = (uint8_t)i;
  }
};

inline const ScanTables& scan_tables() {
  static const ScanTables t;
  return t;
}

}  // namespace hc
