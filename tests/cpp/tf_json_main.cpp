// Loads one camera's transform from a calibration JSON with pcs_b200::load_transform and prints the
// 16 floats as hex bit patterns (tests/test_calibration.py compares them with numpy's float32).
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "pcs_b200_shim.hpp"

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    float tf[16];
    if (!pcs_b200::load_transform(argv[1], argv[2], tf)) {
        std::printf("NOT FOUND\n");
        return 1;
    }
    for (int i = 0; i < 16; ++i) {
        uint32_t bits;
        std::memcpy(&bits, &tf[i], 4);
        std::printf("%08x\n", bits);
    }
    return 0;
}
