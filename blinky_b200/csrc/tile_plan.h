// Tile plan: how the packed lensmap is re-laid out for the tiled warp kernel.
//
// The screen is cut into TW x TH pixel tiles.  At lensmap-build time (host,
// once) each tile is classified:
//   EMPTY   no pixel of the tile is mapped           -> background copy only
//   BOX     every mapped pixel reads one plate and the source texels fit a
//           box of at most 4096 bytes                 -> the kernel stages that
//           box in shared memory with one TMA tensor load and gathers from
//           there; the tile's lensmap entries shrink to 16 bits (offset inside
//           the box, tint, valid)
//   GATHER  anything else (plate seams, singular points, strong minification)
//           -> 32-bit entries, direct global gather
// Entries are stored tile by tile so that a tile's block is one contiguous,
// fully coalesced read.
//
// This replaces nothing in the reference (its lensmap is a flat array of
// pointers, /root/reference/engine/NQ/fisheye.c:427-430); it is the B200 data
// layout for the same information.
#pragma once

#include <cstdint>
#include <vector>

namespace blinky {

constexpr int kTileW = 32;
constexpr int kTileH = 32;
constexpr int kTilePixels = kTileW * kTileH;
constexpr int kMaxBoxBytes = 4096;

// TILE_BOX_FULL: a BOX tile whose 1024 pixels are all mapped and inside the frame
// (the kernel skips every per-pixel validity / background test for it)
enum TileType : uint8_t { TILE_EMPTY = 0, TILE_BOX = 1, TILE_GATHER = 2, TILE_BOX_FULL = 3 };

// 16-bit entry of a BOX tile
constexpr uint16_t kBoxValid = 0x8000;
constexpr int kBoxTintShift = 12;
constexpr uint16_t kBoxOffsetMask = 0x0FFF;

struct TileDesc {       // 16 bytes, read by the kernel
    uint32_t entry_offset;  // byte offset of the tile's entry block (16-byte aligned)
    int16_t box_x, box_y;   // box origin in plate texel coordinates (may be < 0: TMA zero-fills)
    uint8_t plate;
    uint8_t type;           // TileType
    uint8_t box_w16;        // box width / 16  (1..8)
    uint8_t box_h8;         // box height / 8  (1..32)
    uint16_t px, py;        // tile origin on the screen, pixels
};
static_assert(sizeof(TileDesc) == 16, "TileDesc layout is part of the kernel ABI");

struct TilePlan {
    int width = 0, height = 0, platesize = 0;
    int tiles_x = 0, tiles_y = 0;
    std::vector<TileDesc> tiles;        // BOX / BOX_FULL tiles first (n_box of them), then GATHER and EMPTY tiles
    std::vector<uint8_t> entries;       // all entry blocks, tile-ordered
    std::vector<uint16_t> shapes;       // distinct (w16 << 8 | h8) used by BOX tiles
    int n_empty = 0, n_box = 0, n_gather = 0, n_box_full = 0;  // n_box includes n_box_full
    uint64_t box_bytes = 0;             // sum of staged box sizes (bytes per frame through TMA)
};

// packed: [height][width] entries in the BLINKY_LM_* format.  allow_box = false
// forces every non-empty tile to GATHER (e.g. platesize not a multiple of 16,
// which TMA cannot address).
// Tile rows are classified on `threads` host threads; the result does not depend on the thread count.
TilePlan make_tile_plan(const uint32_t *packed, int width, int height, int platesize, bool allow_box, int threads = 1);

}  // namespace blinky
