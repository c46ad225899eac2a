// BekkerSystem.h — Bekker-tire system adapter and its factory (reference src/BekkerSystem.h:8-49); 5 parameters,
// forward only (no adjoint: train_bekker is out of scope).
#pragma once
#include "GpuSystem.h"

template <typename Scalar> class BekkerSystem : public avs::GpuSystem<Scalar, AVS_TIRE_BEKKER> {
public:
    explicit BekkerSystem(const std::shared_ptr<const TerrainMap<Scalar>>& map) : avs::GpuSystem<Scalar, AVS_TIRE_BEKKER>(map) {}
};

template <typename Scalar> class BekkerSystemFactory : public SystemFactory<Scalar> {
public:
    explicit BekkerSystemFactory(const std::shared_ptr<const TerrainMap<Scalar>>& map) : SystemFactory<Scalar>(map) {}
    std::shared_ptr<System<Scalar>> makeSystem() override { return std::make_shared<BekkerSystem<Scalar>>(this->m_terrain_map); }
};
