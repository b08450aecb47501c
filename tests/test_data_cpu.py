"""Native input pipeline (csrc/runtime/data_loader.cc, tepdist_b200/data): windows, labels, determinism, rank sharding, mixture
weights, resume, prefetch ring under a slow consumer."""
import time

import numpy as np
import pytest
import torch

from tepdist_b200.data import TokenLoader, write_token_file


def _files(tmp_path, n=2, length=5000):
    out = []
    for i in range(n):
        # file i holds the arithmetic stream i * 10000 + position: a window's first token identifies (file, offset)
        p, bpt = write_token_file(str(tmp_path / f"shard{i}.bin"), np.arange(length) + i * 10000)
        assert bpt == 2
        out.append(p)
    return out


def test_windows_are_consecutive_tokens_and_labels_are_the_next_token(tmp_path):
    files = _files(tmp_path)
    ld = TokenLoader(files, batch=4, n_ctx=16, seed=3, prefetch=3, steps=5)
    seen = 0
    for feeds in ld:
        tok, lab = feeds["tokens"], feeds["labels"]
        assert tok.shape == (4, 16) and tok.dtype == torch.int32 and lab.shape == (4, 16)
        assert torch.equal(tok[:, 1:], lab[:, :-1])                       # labels = tokens shifted by one
        assert torch.equal(tok[:, 1:] - tok[:, :-1], torch.ones(4, 15, dtype=torch.int32))   # consecutive positions of ONE file
        assert torch.equal(lab[:, -1], tok[:, -1] + 1)
        seen += 1
    assert seen == 5


def test_stream_is_a_pure_function_of_seed_and_step_and_ranks_take_disjoint_rows(tmp_path):
    files = _files(tmp_path)
    whole = [f["tokens"].clone() for f in TokenLoader(files, batch=8, n_ctx=8, seed=7, steps=4)]
    again = [f["tokens"].clone() for f in TokenLoader(files, batch=8, n_ctx=8, seed=7, steps=4, threads=1, prefetch=2)]
    other = [f["tokens"].clone() for f in TokenLoader(files, batch=8, n_ctx=8, seed=8, steps=4)]
    assert all(torch.equal(a, b) for a, b in zip(whole, again))
    assert not all(torch.equal(a, b) for a, b in zip(whole, other))
    # 4 data-parallel ranks of batch 2 each see exactly their rows of the same global batches
    for r in range(4):
        part = [f["tokens"].clone() for f in TokenLoader(files, batch=2, n_ctx=8, seed=7, rank=r, world=4, steps=4)]
        for t in range(4):
            assert torch.equal(part[t], whole[t][2 * r:2 * r + 2]), (r, t)
    # resume: a loader started at step 2 continues the stream
    resumed = [f["tokens"].clone() for f in TokenLoader(files, batch=8, n_ctx=8, seed=7, start_step=2, steps=2)]
    assert torch.equal(resumed[0], whole[2]) and torch.equal(resumed[1], whole[3])


def test_mixture_follows_the_weights_and_int32_files_work(tmp_path):
    a, _ = write_token_file(str(tmp_path / "a.bin"), np.zeros(4000, dtype=np.int64) + 70000, dtype="auto")      # needs int32
    b, _ = write_token_file(str(tmp_path / "b.bin"), np.zeros(4000, dtype=np.int64) + 80000, dtype="int32")
    ld = TokenLoader(mix=[(a, 3.0), ([b], 1.0)], batch=64, n_ctx=4, bytes_per_token=4, steps=16)
    from_a = sum(int((f["tokens"][:, 0] == 70000).sum()) for f in ld)
    frac = from_a / (64 * 16)
    assert 0.70 < frac < 0.80, frac                                       # 3 : 1


def test_synthetic_tokens_stay_below_the_vocabulary(tmp_path):
    ld = TokenLoader(synthetic_vocab=1000, batch=4, n_ctx=32, steps=3)
    for f in ld:
        assert int(f["tokens"].min()) >= 0 and int(f["tokens"].max()) < 1000 and torch.equal(f["tokens"][:, 1:], f["labels"][:, :-1])
    with pytest.raises(ValueError):
        TokenLoader(batch=1, n_ctx=4)


def test_prefetch_ring_runs_ahead_of_a_slow_consumer_but_never_overwrites_the_batch_in_use(tmp_path):
    files = _files(tmp_path)
    ld = TokenLoader(files, batch=2, n_ctx=8, seed=1, prefetch=3, threads=2, steps=6)
    ref = [f["tokens"].clone() for f in TokenLoader(files, batch=2, n_ctx=8, seed=1, steps=6)]
    first = next(ld)
    time.sleep(0.3)                                                       # the workers fill every free slot meanwhile ...
    assert ld._loader.batches_filled() == 3                               # ... all 3 slots, and then wait: slot 0 is in use
    assert torch.equal(first["tokens"], ref[0])                           # the held batch was not overwritten
    rest = [f["tokens"].clone() for f in ld]
    assert len(rest) == 5 and all(torch.equal(a, b) for a, b in zip(rest, ref[1:]))


def test_a_file_shorter_than_one_window_is_an_error(tmp_path):
    p, _ = write_token_file(str(tmp_path / "tiny.bin"), np.arange(5))
    ld = TokenLoader(p, batch=1, n_ctx=4, steps=1)                        # 5 tokens = exactly one window of n_ctx + 1
    assert torch.equal(next(ld)["tokens"][0], torch.arange(4, dtype=torch.int32))
    with pytest.raises(RuntimeError):
        TokenLoader(p, batch=1, n_ctx=8).sample(0)
    with pytest.raises(RuntimeError, match="shorter than one window"):      # raised on a worker thread, re-raised to the consumer
        next(TokenLoader(p, batch=1, n_ctx=8))
