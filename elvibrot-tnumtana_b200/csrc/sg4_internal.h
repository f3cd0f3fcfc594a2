// sg4_internal.h -- shared declarations of the evr_sg4 library (not part of the C-ABI).
#pragma once
#include <cstdint>
#include <string>

#define EVR_MAXD 32          // max number of SG4 modes (HH 21-D is the largest shipped input)
#define EVR_MAXCH 4          // max nb0 (electronic channels) handled in registers

namespace evr {
int fail(const std::string &msg);      // records the message for evr_sg4_last_error(), returns 1
}
