// types/SystemFactory.h — factory handed to Trainer (reference src/types/SystemFactory.h:9-17).
#pragma once
#include <memory>
#include "types/System.h"
#include "types/TerrainMap.h"

template <typename Scalar> class SystemFactory {
public:
    typedef std::shared_ptr<const TerrainMap<Scalar>> MapPtr;
    explicit SystemFactory(const MapPtr& map) : m_terrain_map{map} {}
    virtual ~SystemFactory() = default;
    virtual std::shared_ptr<System<Scalar>> makeSystem() = 0;
    const MapPtr& terrainMap() const { return m_terrain_map; }  // extension: the multi-GPU Trainer configures one context per GPU

protected:
    const MapPtr m_terrain_map;
};
