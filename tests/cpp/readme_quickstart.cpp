// The reference's README quick start (README.md:51-68) through the C++ mirror of its API.
// Built by tests/test_abi.py (compile + link check on CPU) and run by tests/test_gpu_cpp_mirror.py.
#include <cstdio>

#include "voxelis_b200.hpp"

int main() {
    using namespace voxelis;
    try {
        auto interner = VoxInterner<uint8_t>::with_memory_budget(256u << 20);  // "256 MB budget"
        VoxTree<uint8_t> tree(5);                                               // 32^3 voxels
        auto batch = tree.create_batch();
        batch.fill(interner, 2);
        batch.set(interner, {1, 2, 3}, 7);
        if (!tree.apply_batch(interner, batch)) return 2;
        auto a = tree.get(interner, {1, 2, 3});
        auto b = tree.get(interner, {0, 0, 0});
        if (!a || *a != 7 || !b || *b != 2) return 3;
        if (tree.is_leaf() || tree.is_empty() || interner.get_ref(tree.get_root_id()) != 1) return 4;
        auto dense = tree.to_vec(interner);
        size_t sevens = 0;
        for (auto v : dense) sevens += v == 7;
        if (sevens != 1 || dense.size() != 32u * 32 * 32) return 5;
        tree.clear(interner);                                                   // releases every node
        if (!tree.is_empty()) return 6;
        std::printf("ok\n");
        return 0;
    } catch (const Error& e) {
        std::printf("error %d: %s\n", e.code, e.what());
        return 1;
    }
}
