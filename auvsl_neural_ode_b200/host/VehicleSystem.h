// VehicleSystem.h — NN-tire system adapter and its factory (reference src/VehicleSystem.h:13-51).
#pragma once
#include "GpuSystem.h"

template <typename Scalar> class VehicleSystem : public avs::GpuSystem<Scalar, AVS_TIRE_NN> {
public:
    explicit VehicleSystem(const std::shared_ptr<const TerrainMap<Scalar>>& map) : avs::GpuSystem<Scalar, AVS_TIRE_NN>(map) {}
};

template <typename Scalar> class VehicleSystemFactory : public SystemFactory<Scalar> {
public:
    explicit VehicleSystemFactory(const std::shared_ptr<const TerrainMap<Scalar>>& map) : SystemFactory<Scalar>(map) {}
    std::shared_ptr<System<Scalar>> makeSystem() override { return std::make_shared<VehicleSystem<Scalar>>(this->m_terrain_map); }
};
