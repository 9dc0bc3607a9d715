"""Host-side logic of the CRD mirror that needs no GPU: which formulation of the step a ContrastMemory selects
(``_step_variant``), shard bookkeeping, and the state_dict surface the reference's checkpoints rely on."""
import pytest
import torch

STREAM = 0x200


def test_step_variant_selection_rules(pkg):
    """streaming=None (default): bf16 banks take the bank-streaming tensor-core step when the step draws >= 2 samples per
    resident row, feat_dim is 128 and the batch fits 48 anchors; fp32 banks never do; True forces or raises."""
    bf = pkg.ContrastMemory(128, 8192, 4096, 0.07, 0.5, bank_dtype=torch.bfloat16)
    assert bf.streaming is None and bf.STREAM == STREAM
    assert bf._step_variant(46, 4097, 128) & STREAM            # 188 K samples over 8 K rows
    assert not (bf._step_variant(2, 4097, 128) & STREAM)       # 8 K samples: gathering is cheaper than streaming the bank
    assert not (bf._step_variant(49, 4097, 128) & STREAM)      # unsupported batch: automatic mode falls back to the gather step
    assert not (bf._step_variant(46, 4097, 256) & STREAM)      # unsupported width
    bf.streaming = False
    assert not (bf._step_variant(46, 4097, 128) & STREAM)
    bf.streaming = True
    assert bf._step_variant(2, 4097, 128) & STREAM             # forced: no profitability test
    with pytest.raises(RuntimeError):
        bf._step_variant(49, 4097, 128)
    f32 = pkg.ContrastMemory(128, 8192, 4096, 0.07, 0.5)
    assert not (f32._step_variant(46, 4097, 128) & STREAM)
    f32.variant = STREAM                                        # an explicit variant always wins
    assert f32._step_variant(46, 4097, 128) == STREAM


def test_step_variant_counts_only_the_samples_of_the_shard(pkg):
    """A shard holding 1/8 of the rows sees 1/8 of a replicated contrast_idx, but all K+1 columns when every rank draws its
    own in-shard negatives (k_total > 0)."""
    N, K = 65536, 4096
    sh = pkg.ContrastMemory(128, N, K, 0.07, 0.5, row_begin=0, row_end=N // 8, bank_dtype=torch.bfloat16)
    assert sh.memory_v1.shape == (N // 8, 128)
    # replicated index list: 46 * 4097 / 8 = 23.5 K samples on 8 K rows -> 2.9 per row: streams
    assert sh._step_variant(46, K + 1, 128) & STREAM
    # ... 4 anchors: 2 K samples on 8 K rows: gathers
    assert not (sh._step_variant(4, K + 1, 128) & STREAM)
    sh.k_total = 8 * K                                         # local negatives: every column lands on this shard
    assert sh._step_variant(4, K + 1, 128) & STREAM            # 16 K samples on 8 K rows


def test_state_dict_surface_and_bank_layout(pkg):
    """The banks are buffers named as in the published module, whatever the internal layout; interleaved banks are two
    views of one [rows, 2, D] allocation (bank2 = bank1 + D elements, pitch 2D)."""
    mem = pkg.ContrastMemory(128, 1000, 64, 0.07, 0.5)
    sd = mem.state_dict()
    assert set(sd) >= {"params", "memory_v1", "memory_v2"}
    assert sd["memory_v1"].shape == (1000, 128) and sd["memory_v2"].shape == (1000, 128)
    assert sd["params"].tolist()[:2] == [64.0, pytest.approx(0.07)]
    m1, m2 = mem.memory_v1, mem.memory_v2
    assert m1.stride() == (256, 1) and m2.stride() == (256, 1)
    assert m2.data_ptr() - m1.data_ptr() == 128 * m1.element_size()
    dense = pkg.ContrastMemory(128, 1000, 64, 0.07, 0.5, interleave=False)
    assert dense.memory_v1.is_contiguous() and dense.memory_v2.is_contiguous()
    # a checkpoint written by one layout loads into the other
    dense.load_state_dict(sd)
    assert torch.equal(dense.memory_v1, m1) and torch.equal(dense.memory_v2, m2)


def test_sampler_state_round_trip(pkg):
    """The negative sampler's (seed, offset) is exposed for checkpointing without touching the state_dict keys."""
    mem = pkg.ContrastMemory(128, 1000, 15, seed=77)
    mem.multinomial.offset = 12345
    st = mem.sampler_state()
    assert st == {"seed": 77, "offset": 12345}
    other = pkg.ContrastMemory(128, 1000, 15, seed=1)
    other.load_sampler_state(st)
    assert other.sampler_state() == st
    assert not any("extra_state" in k for k in mem.state_dict())


def test_sweep_selection_rules(pkg):
    """Band-sorted lists are chosen for shards of about the L2's size and larger whose rows repeat within the step, never for
    small banks, few samples, in-L2 shards, and are forced / forbidden by ``sweep``; in-shard negatives keep the hint bit."""
    mem = pkg.ContrastMemory(128, 1000, 15)

    def variant(rows, n_data, B, K1, k_total=0, sweep=None, base=0):
        mem.row_begin, mem.row_end, mem.nLem, mem.k_total, mem.sweep, mem.variant = 0, rows, n_data, k_total, sweep, base
        return mem._step_variant(B, K1, 128)

    S = mem.SWEEP
    assert variant(1_000_000, 1_000_000, 46, 65537) & S                 # headline: 1 GB bank, 3 draws per row
    assert variant(500_000, 1_000_000, 46, 65537) & S                   # 2-way shard
    assert variant(125_000, 1_000_000, 46, 65537) & S                   # 8-way shard: 128 MB
    assert not variant(90_000, 90_000, 46, 16385) & S                   # config 0: 92 MB bank inside the L2
    assert not variant(1_000_000, 1_000_000, 46, 4097) & S              # 188 k samples: no repeats to catch
    assert not variant(1_000_000, 1_000_000, 46, 65537, sweep=False) & S
    assert variant(8192, 8192, 8, 256, sweep=True) & S
    assert not variant(1_000_000, 1_000_000, 46, 65537, base=0x40) & S  # an explicit compact request is kept
    assert variant(1_000_000, 8_000_000, 46, 65537, k_total=8 * 65536) & S   # weak scaling: every entry in-shard
