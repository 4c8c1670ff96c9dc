// tile_engine.cuh — placeholder until the shared-memory-resident engine lands (next commit).
#pragma once
#include <string>

#include "formula.hpp"

namespace odesat {

template <typename T> struct TileEngine {
    static bool supports(const odesat_formula&, int64_t, std::string* why) {
        if (why) *why = "tile engine not built yet";
        return false;
    }
    static bool preferred(const odesat_formula&, int64_t) { return false; }
    TileEngine(const odesat_formula&, int64_t, int, cudaStream_t, int64_t*) {}
    void reset_control() {}
    int64_t import_state(const T*, const T*, const T*, int64_t) { return 0; }
    int64_t export_state(T*, T*, T*, int64_t) { return 0; }
    int64_t run_fixed(T, T, int64_t, int, int32_t*, int64_t) { return 0; }
};

}  // namespace odesat
