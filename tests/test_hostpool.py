"""CPU: the host worker pool of the library (csrc/hostpool.hpp) -- every task runs exactly once, groups are independent,
wait() returns only when its group is done (the caller helps), a pool without workers runs tasks inline."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include "hostpool.hpp"
#include <cstdio>
int main() {
    dbslmm::HostPool pool;
    pool.start(6);
    pool.start(4);                                 // start() only grows
    if (pool.size() != 6) { std::printf("size %d\n", pool.size()); return 1; }
    long total = 0;
    for (int rep = 0; rep < 300; ++rep) {
        dbslmm::TaskGroup a, b;
        std::atomic<int> na{0}, nb{0};
        int slots[24] = {0};
        for (int i = 0; i < 24; ++i) pool.submit(a, [&na, &slots, i]() { long s = 0; for (int k = 0; k < 20000; ++k) s += k ^ i; slots[i] += 1 + (int)(s & 0); ++na; });
        for (int i = 0; i < 7; ++i) pool.submit(b, [&nb]() { ++nb; });
        pool.wait(b);
        if (nb != 7 || b.pending != 0) { std::printf("group b: %d\n", (int)nb); return 2; }
        pool.wait(a);
        if (na != 24 || a.pending != 0) { std::printf("group a: %d\n", (int)na); return 3; }
        for (int i = 0; i < 24; ++i) if (slots[i] != 1) { std::printf("slot %d ran %d times\n", i, slots[i]); return 4; }
        total += na + nb;
        if (rep % 100 == 0) std::this_thread::sleep_for(std::chrono::milliseconds(1));      // let the workers go to sleep
    }
    dbslmm::HostPool none;
    dbslmm::TaskGroup g;
    int x = 0;
    none.submit(g, [&x]() { x = 5; });
    none.wait(g);
    if (x != 5) return 5;
    std::printf("ok %ld\n", total);
    return 0;
}
'''


def test_host_pool(tmp_path):
    src = tmp_path / "pool_test.cpp"
    src.write_text(SRC)
    exe = tmp_path / "pool_test"
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "dbslmm_b200", "csrc"), str(src), "-o", str(exe)],
                   check=True)
    out = subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True, timeout=120)
    assert out.returncode == 0, out.stdout
    assert out.stdout.strip() == "ok 9300"
