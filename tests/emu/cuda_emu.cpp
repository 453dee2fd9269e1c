// cuda_emu.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_emu.h).
#include "cuda_emu.h"

namespace emu {
thread_local BlockCtx* tls_ctx = nullptr;
thread_local uint3 tls_threadIdx, tls_blockIdx;
thread_local dim3 tls_blockDim, tls_gridDim;

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  const unsigned nt = block.x;
  std::vector<unsigned char> dyn(smem + 64);
  unsigned char* dyn_aligned = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dyn.data()) + 63) & ~uintptr_t(63));
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        std::barrier<> bar(nt);
        std::vector<uint64_t> shfl(nt);
        std::vector<std::unique_ptr<std::barrier<>>> wbar;
        for (unsigned w = 0; w < (nt + 31) / 32; ++w) {
          unsigned lanes = std::min(32u, nt - w * 32);
          wbar.emplace_back(new std::barrier<>(lanes));
        }
        BlockCtx ctx{&bar, dyn_aligned, &shfl, &wbar};
        std::vector<std::thread> th;
        th.reserve(nt);
        for (unsigned t = 0; t < nt; ++t) {
          th.emplace_back([&, t]() {
            tls_ctx = &ctx;
            tls_threadIdx = uint3{t, 0, 0};
            tls_blockIdx = uint3{bx, by, bz};
            tls_blockDim = block;
            tls_gridDim = grid;
            body();
          });
        }
        for (auto& x : th) x.join();
      }
}
}  // namespace emu
