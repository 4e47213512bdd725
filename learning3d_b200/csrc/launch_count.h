// Process-wide count of kernel launches issued by libl3d_b200 (bench.py's `gpu_launches`).
#pragma once
#include <stdint.h>
namespace l3d {
void count_launch(int n = 1);
}
