"""The five BASELINE.json configurations as parity cases (CUDA path through the C ABI vs the oracle):
configs[0] at its full size, the others scaled so that the oracle finishes in seconds.  The inputs are
generated on the device by the deterministic generators of include/modgpu_synth.h (the host build of
the same header is checked against them in test_gpu_parity.py) and copied back for the oracle."""
import numpy as np
import pytest

import hostemul as he

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mg():
    import modimizer_b200 as m
    m.require_device()
    return m


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    return torch


def device_reads(torch, spec, n_reads, ont=False):
    """(device tensor, host copy) of n_reads synthetic reads"""
    from modimizer_b200 import synth
    n = n_reads * spec.readLen
    buf = torch.empty(n + 64, dtype=torch.uint8, device="cuda:0")
    synth.reads_device(spec, 0, n_reads, ont, buf.data_ptr())
    torch.cuda.synchronize()
    return buf, buf[:n].cpu().numpy()


def device_genome(torch, seed, n, dup_mode):
    from modimizer_b200 import synth
    buf = torch.empty(n + 64, dtype=torch.uint8, device="cuda:0")
    synth.genome_device(seed, 0, n, dup_mode, buf.data_ptr())
    torch.cuda.synchronize()
    return buf, buf[:n].cpu().numpy()


def check_same_modset(ms, orc, oms):
    gv, gd, gi = ms.sorted_dump()
    ov, od, oi = orc.modset_sorted(oms)
    assert ms.max == orc._modset_max(oms)
    assert np.array_equal(gv, ov) and np.array_equal(gd, od) and np.array_equal(gi, oi)
    assert np.array_equal(ms.histogram(), orc.modset_hist(oms))
    assert ms.summary() == orc.modset_summary(oms)


def test_config0_modutils_build_count_full_size(mg, torch_cuda, orc):
    """configs[0]: 10 Mb genome, 30 000 x 10 kb reads (30x), k=19 d=31, tableBits 24 - FULL SIZE, device path
    and host path; sorted dump, -H histogram, -s 10 45 75 classes and the summary text identical"""
    sp = he.read_spec(12345, 10_000_000, 7, 10_000)
    nreads = 30_000
    d_reads, reads = device_reads(torch_cuda, sp, nreads)
    offs = np.arange(nreads + 1, dtype=np.uint64) * np.uint64(10_000)
    d_offs = torch_cuda.from_numpy(offs.view(np.int64)).cuda()
    oms = orc.modset_new(24, 19, 31, 17)
    otot = orc.modset_add(oms, reads, offs)
    ms = mg.Modset(24, 19, 31, 17)
    ms2 = mg.Modset(24, 19, 31, 17)
    try:
        assert ms.add_device(d_reads.data_ptr(), d_offs.data_ptr(), nreads, len(reads)) == otot
        check_same_modset(ms, orc, oms)
        assert ms2.add(reads, offs, is_ascii=0) == otot          # pinned staging, chunked, pipelined
        check_same_modset(ms2, orc, oms)
        # about 30x: the modal depth bin sits near 30 and nearly every k-mer of the genome was seen
        h = ms.histogram()
        assert 25 <= int(np.argmax(h[1:]) + 1) <= 35
        assert abs(ms.max - (10_000_000 - 18) / 31) < 0.01 * 10_000_000 / 31
        c = ms.set_copy(10, 45, 75)
        orc._modset_setcopy(oms, 10, 45, 75)
        oi = orc.modset_sorted(oms)[2]
        assert list(c) == [int((oi == j).sum()) for j in range(4)] and c[1] > 0.95 * ms.max
        assert np.array_equal(ms.sorted_dump()[2], oi)
        assert ms.summary() == orc.modset_summary(oms)
    finally:
        ms.close(); ms2.close(); orc._modset_free(oms)


def test_config2_hifi_reads_count(mg, torch_cuda, orc):
    """configs[2] scaled: HiFi-like 15 kb reads with 0.1 % substitutions, 20x of a 4 Mb genome with planted
    duplications, k=31 d=64 and k=19 d=31; error k-mers show up as the depth-1 tail of the histogram"""
    sp = he.read_spec(2024, 4_000_000, 11, 15_000, sub_ppm=1000, dup_mode=1)
    nreads = 5_400
    d_reads, reads = device_reads(torch_cuda, sp, nreads)
    offs = np.arange(nreads + 1, dtype=np.uint64) * np.uint64(15_000)
    for (k, d, bits) in ((31, 64, 22), (19, 31, 23)):
        oms = orc.modset_new(bits, k, d, 17)
        otot = orc.modset_add(oms, reads, offs)
        ms = mg.Modset(bits, k, d, 17)
        try:
            assert ms.add(reads, offs, is_ascii=0) == otot
            check_same_modset(ms, orc, oms)
            h = ms.histogram()
            assert h[1] > h[5]                                   # error k-mers: a depth-1 peak well above the valley
        finally:
            ms.close(); orc._modset_free(oms)


def test_config3_illumina_pairs_histogram_and_classes(mg, torch_cuda, orc):
    """configs[3] scaled: 2 x 150 bp pairs (mate 2 = reverse complement of the far fragment end), 40x of a 2 Mb
    genome, 0.3 % substitutions, k=19 d=31: half a million short reads, i.e. a read boundary every 150 bases;
    histogram, single-copy classification and summary identical"""
    sp = he.read_spec(778, 2_000_000, 13, 150, sub_ppm=3000, frag_len=400, pair_mode=1, dup_mode=1)
    nreads = 2 * 266_000
    d_reads, reads = device_reads(torch_cuda, sp, nreads)
    offs = np.arange(nreads + 1, dtype=np.uint64) * np.uint64(150)
    oms = orc.modset_new(22, 19, 31, 17)
    otot = orc.modset_add(oms, reads, offs)
    ms = mg.Modset(22, 19, 31, 17)
    try:
        d_offs = torch_cuda.from_numpy(offs.view(np.int64)).cuda()
        assert ms.add_device(d_reads.data_ptr(), d_offs.data_ptr(), nreads, len(reads)) == otot
        check_same_modset(ms, orc, oms)
        # thresholds from the histogram: effective depth 40 x (150-18)/150 = 35
        c = ms.set_copy(10, 60, 100)
        orc._modset_setcopy(oms, 10, 60, 100)
        oi = orc.modset_sorted(oms)[2]
        assert list(c) == [int((oi == j).sum()) for j in range(4)]
        assert c[0] > 0 and c[1] > 0 and c[2] > 0               # errors, single copy, planted duplicates
        assert np.array_equal(ms.sorted_dump()[2], oi)
        assert ms.summary() == orc.modset_summary(oms)
    finally:
        ms.close(); orc._modset_free(oms)


def test_config1_and_4_modmap_index_and_ont_queries(mg, torch_cuda, orc):
    """configs[1] and [4] scaled: modmap index of a 24-record 12 Mb genome with duplications (k=31 d=64,
    tableBits 24) - every Reference array identical - then 1 500 ONT-like 10 kb reads with 10 % errors
    (substitutions, insertions, deletions) matched against it: seeds, Q counters and hit lists identical"""
    glen = 12_000_000
    _, genome = device_genome(torch_cuda, 4242, glen, 1)
    cuts = (np.arange(25, dtype=np.float64) * (glen / 24)).astype(np.uint64)
    cuts[-1] = glen
    R = mg.Reference(24, 31, 64, 17, genome, cuts, is_ascii=0)
    oref, ocounts = orc.ref_build(24, 31, 64, 17, genome, cuts)
    try:
        assert list(R.counts) == list(ocounts)
        assert ocounts[1] > 0 and ocounts[2] > 0
        g = R.export(); o = orc.ref_export(oref)
        for key in ("index", "offset", "id", "depth", "loc", "rev"):
            assert np.array_equal(g[key], o[key]), key
        sp = he.read_spec(4242, glen, 5, 10_000, 30_000, 30_000, 40_000, dup_mode=1)
        nreads = 1_500
        _, reads = device_reads(torch_cuda, sp, nreads, ont=True)
        roffs = np.arange(nreads + 1, dtype=np.uint64) * np.uint64(10_000)
        gq = R.query(reads, roffs, is_ascii=0)
        oq = orc.ref_query(oref, reads, roffs)
        for key in ("seedOff", "index", "pos", "hitId", "hitOffset", "counters"):
            assert np.array_equal(gq[key], oq[key]), key
        cnt = oq["counters"].sum(axis=0)
        assert cnt[1] > 0 and cnt[0] > cnt[1]                    # 10 % errors: most 31-mers miss, some single-copy hits
    finally:
        R.close(); orc._ref_free(oref)
