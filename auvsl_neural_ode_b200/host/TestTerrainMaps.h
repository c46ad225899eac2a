// TestTerrainMaps.h — Flat / Slope / Bumpy maps (reference src/TestTerrainMaps.{h,cpp}) plus a gridded map
// (extension: bilinear height field with optional per-cell Bekker soil parameters).
#pragma once
#include <cmath>
#include "types/TerrainMap.h"

template <typename Scalar> class FlatTerrainMap : public TerrainMap<Scalar> {
public:
    Scalar getAltitude(const Scalar&, const Scalar&) const override { return Scalar(0.0); }
    avs::TerrainKind deviceKind() const override { return avs::TERRAIN_FLAT; }
};

template <typename Scalar> class SlopeTerrainMap : public TerrainMap<Scalar> {
public:
    Scalar getAltitude(const Scalar& x, const Scalar&) const override { return 0.1 * x; }
    avs::TerrainKind deviceKind() const override { return avs::TERRAIN_SLOPE; }
};

// one 0.2 m sine bump on 2 < x < 3.6, flat elsewhere (a window, not a clamp: TestTerrainMaps.cpp:18-28)
template <typename Scalar> class BumpyTerrainMap : public TerrainMap<Scalar> {
public:
    Scalar getAltitude(const Scalar& x, const Scalar&) const override {
        Scalar a = x - 2.0;
        if (!(a < 1.6)) a = 0.0;
        if (!(a > 0.0)) a = 0.0;
        const Scalar h = .2 * std::sin(a);
        return h > 0.0 ? h : Scalar(0.0);
    }
    avs::TerrainKind deviceKind() const override { return avs::TERRAIN_BUMPY; }
};

template <typename Scalar> class GridTerrainMap : public TerrainMap<Scalar> {
public:
    explicit GridTerrainMap(avs::TerrainGrid g) : g_(std::move(g)) {}
    Scalar getAltitude(const Scalar& x, const Scalar& y) const override {
        double u = (x - g_.x0) / g_.dx, v = (y - g_.y0) / g_.dy;
        u = u < 0 ? 0 : (u > g_.nx - 1 ? g_.nx - 1 : u);
        v = v < 0 ? 0 : (v > g_.ny - 1 ? g_.ny - 1 : v);
        int i = (int)std::floor(u), j = (int)std::floor(v);
        if (i > g_.nx - 2) i = g_.nx - 2;
        if (j > g_.ny - 2) j = g_.ny - 2;
        const double fu = u - i, fv = v - j;
        const double* h = g_.height.data();
        const double a = h[j * g_.nx + i] + fu * (h[j * g_.nx + i + 1] - h[j * g_.nx + i]);
        const double b = h[(j + 1) * g_.nx + i] + fu * (h[(j + 1) * g_.nx + i + 1] - h[(j + 1) * g_.nx + i]);
        return a + fv * (b - a);
    }
    avs::TerrainKind deviceKind() const override { return avs::TERRAIN_GRID; }
    const avs::TerrainGrid* deviceGrid() const override { return &g_; }
private:
    avs::TerrainGrid g_;
};
