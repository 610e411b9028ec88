/* version.h — version of the afQuantumSim API this host layer mirrors
 * (reference include/version.h:21-49) and of the engine beneath it. */
#pragma once
#include <string>

#define AQS_VERSION_MAJOR 1
#define AQS_VERSION_MINOR 0
#define AQS_VERSION_PATCH 0
#define AQS_VERSION "1.0.0"
#define AQS_B200_ENGINE_VERSION "0.1.0"

namespace aqs {
inline std::string get_version() { return AQS_VERSION; }
inline std::string get_engine_version() { return AQS_B200_ENGINE_VERSION; }
}  // namespace aqs
