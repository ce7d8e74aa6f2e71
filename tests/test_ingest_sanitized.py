"""CPU: the C++ host ingest of the `dbslmm` command line (dbslmm_b200/host/ingest.cpp) compiled with AddressSanitizer and
UndefinedBehaviorSanitizer, run on the C1 fixture files: same SNP / block decisions as the Python mirror (hostio.py, itself
pinned to the reference's counts in test_hostio.py), and no sanitizer report."""
import os
import subprocess

import numpy as np

from dbslmm_b200 import hostio as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

MAIN = r'''
#include "ingest.hpp"
#include <cstdio>
using namespace dbslmm_host;
int main(int argc, char** argv) {
    const std::string d = argv[1];
    const int n_ref = get_row(d + "/ref.fam");
    BimMap bim;
    const int64_t nsnp = read_bim(d + "/ref.bim", bim);
    std::vector<uint8_t> bed;
    if (!read_bed(d + "/ref.bed", nsnp, n_ref, bed)) return 2;
    std::vector<uint8_t> bed2(bed.size());
    if (!read_bed_into(d + "/ref.bed", nsnp, n_ref, bed2.data()) || bed2 != bed) return 3;
    Summ summ;
    if (!read_summ(d + "/summ.txt", summ)) return 4;
    std::vector<Block> blocks;
    if (!read_block(d + "/blocks.bed", blocks)) return 5;
    Info inter, info;
    std::vector<char> matched;
    int dis = 0, mafc = 0;
    match_ref(summ, bim, nullptr, 1.0, inter, matched, dis, mafc);
    const int n_in = add_block(inter, blocks, info);
    const std::vector<int32_t> off = block_offsets(info, (int)blocks.size());
    unsigned long long bedsum = 0;
    for (uint8_t b : bed) bedsum = bedsum * 1315423911ull + b;
    long long possum = 0;
    for (size_t i = 0; i < info.size(); ++i) possum += (long long)info.pos[i] * (long long)(i + 1);
    std::printf("%d %lld %zu %zu %zu %d %d %d %zu %d %llu %lld\n", n_ref, (long long)nsnp, summ.size(), blocks.size(), inter.size(), dis, mafc,
                n_in, info.size(), off.back(), bedsum, possum);
    return 0;
}
'''


def test_cpp_ingest_under_asan_ubsan(tmp_path):
    d = np.load(os.path.join(GOLD, "c1_testdat.npz"))
    (tmp_path / "ref.bim").write_text(str(d["bim_txt"]))
    (tmp_path / "ref.fam").write_text("x\n" * int(d["fam_lines"]))
    with open(tmp_path / "ref.bed", "wb") as f:
        f.write(bytes([0x6C, 0x1B, 0x01]) + d["bed"].tobytes())
    (tmp_path / "summ.txt").write_text(str(d["summary_txt"]))
    (tmp_path / "blocks.bed").write_text(str(d["block_txt"]))
    src = tmp_path / "ingest_main.cpp"
    src.write_text(MAIN)
    exe = tmp_path / "ingest_san"
    host = os.path.join(ROOT, "dbslmm_b200", "host")
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-pthread", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
                    "-I", host, str(src), os.path.join(host, "ingest.cpp"), "-o", str(exe)], check=True)
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1")
    out = subprocess.run([str(exe), str(tmp_path)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, timeout=120)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "ERROR" not in out.stderr and "runtime error" not in out.stderr, out.stderr[-2000:]
    n_ref, nsnp, n_summ, n_blocks, n_inter, dis, mafc, n_in, n_info, tot, bedsum, possum = [int(x) for x in out.stdout.split()]
    # the Python mirror on the same files
    bim, nsnp_py = H.read_bim(tmp_path / "ref.bim")
    summ = H.read_summ(tmp_path / "summ.txt")
    keep, pos = H.match_ref(summ, bim, None, 1.0)
    bs, be = H.read_block(tmp_path / "blocks.bed")
    blk = H.add_block(summ.ps[keep], bs, be)
    assert (n_ref, nsnp) == (H.read_fam_count(tmp_path / "ref.fam"), nsnp_py) == (400, 723)
    assert n_summ == len(summ.snp) == 996 and n_blocks == bs.size == 133
    assert n_inter == keep.size == 717
    assert n_info == tot == int((blk >= 0).sum()) and n_in == 133          # (add_block returns the block count, like SNPPROC::addBlock)
    h = 0
    for b in d["bed"].tobytes():
        h = (h * 1315423911 + b) % (1 << 64)
    assert bedsum == h
    order = np.argsort(blk[blk >= 0], kind="stable")
    assert possum == int(np.sum(pos[blk >= 0][order].astype(np.int64) * np.arange(1, tot + 1)))
