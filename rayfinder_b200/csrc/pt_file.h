// In-memory form of a .pt scene container (nlrs::PtFormat, pt-format/pt_format.hpp:18-43), shared by the codec
// (host_pt_format.cpp) and the scene baker (host_baker.cpp).
#pragma once

#include "rf_internal.h"

#include <cstdint>
#include <vector>

namespace rfb200
{
// bytes per element of the 13 arrays, in the order of the RF_PT_* enum (the C++ structs the reference memcpy's)
constexpr std::uint64_t RF_PT_ELEM_SIZE[RF_PT_NUM_ARRAYS] = {48, 36, 48, 80, 16, 16, 8, 4, 16, 16, 16, 16, 4};

struct PtTextureData
{
    std::uint32_t              width = 0, height = 0;
    std::vector<std::uint32_t> pixels; // BGRA8
};
} // namespace rfb200

struct rf_pt_file
{
    std::vector<std::uint8_t>          arrays[RF_PT_NUM_ARRAYS];
    std::vector<rfb200::PtTextureData> textures;

    std::uint64_t count(int which) const { return arrays[which].size() / rfb200::RF_PT_ELEM_SIZE[which]; }
};
