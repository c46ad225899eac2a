// types/TerrainMap.h — terrain interface of the drop-in host API (reference src/types/TerrainMap.h:6-11).
// getAltitude() keeps its signature for host-side callers.  The GPU kernels cannot call a host virtual, so a
// map additionally describes itself: one of the analytic kinds the kernels implement, or a sampled grid.
#pragma once
#include <memory>
#include <vector>

namespace avs {
enum TerrainKind { TERRAIN_FLAT = 0, TERRAIN_SLOPE = 1, TERRAIN_BUMPY = 2, TERRAIN_GRID = 3, TERRAIN_CUSTOM = -1 };
struct TerrainGrid {
    std::vector<double> height;  // [ny][nx]
    std::vector<double> soil;    // [5][ny][nx] or empty
    int nx = 0, ny = 0;
    double x0 = 0, y0 = 0, dx = 1, dy = 1;
};
}  // namespace avs

template <typename Scalar> class TerrainMap {
public:
    virtual ~TerrainMap() = default;
    virtual Scalar getAltitude(const Scalar& x, const Scalar& y) const = 0;
    // device description; user-defined maps default to CUSTOM and are sampled onto a grid by the systems
    virtual avs::TerrainKind deviceKind() const { return avs::TERRAIN_CUSTOM; }
    virtual const avs::TerrainGrid* deviceGrid() const { return nullptr; }
};
