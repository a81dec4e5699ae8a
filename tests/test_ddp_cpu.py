"""CPU, world_size 2 over gloo: the bucket planner and the overlapped gradient reducer of
bifold_b200.ddp (the N>1 path of SURVEY.md §8e).  The CUDA kernels are not involved here."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bifold_b200.ddp import GradReducer, plan_groups

NAMES = ["siglip_model.model.text_model.encoder.layers.0.self_attn.q_proj.lora_A.siglip_adapter.weight",
         "siglip_model.model.vision_model.encoder.layers.0.self_attn.q_proj.lora_B.siglip_adapter.weight",
         "text_token", "image_token", "pick_place.fusion.token_type_embeddings.weight",
         "pick_place.fusion.transformer_encoder.layers.0.0.norm.weight",
         "pick_place.fusion.transformer_encoder.layers.0.0.fn.to_qkv.weight",
         "pick_place.fusion.transformer_encoder.layers.1.0.norm.weight",
         "pick_place.left_pick_decoder.decoder_net.0.weight", "pick_place.left_pick_decoder.decoder_net.0.bias",
         "context_pos_embedding"]
SIZES = [8 * 64, 64 * 8, 64, 64, 128, 64, 3 * 64 * 64, 64, 32 * 64, 30, 5 * 64]


def test_plan_groups_covers_arena_once():
    groups = plan_groups(NAMES, SIZES)
    assert set(groups) == {"text", "vision", "tokens", "fusion.0", "fusion.1", "decoders"}
    total = sum((s + 7) // 8 * 8 for s in SIZES)
    covered = sorted(r for rs in groups.values() for r in rs)
    pos = 0
    for off, n in covered:
        assert off == pos
        pos += n
    assert pos == total
    assert len(groups["tokens"]) == 2          # tokens + type-embedding, then context_pos_embedding at the end
    assert len(groups["decoders"]) == 1        # weight and (padded) bias merged into one range


def test_bucketed_reducer_merges_ranges_and_flushes(monkeypatch):
    """GradReducer bucketing (single process, all_reduce stubbed): per-group mode issues one collective per arena range; with a
    bucket size the announced ranges are collected, adjacent ones merge, and the bucket goes out when it is full, when the fusion
    part ends ("tokens") and at wait() - every arena element is reduced exactly once either way."""
    calls = []

    class _Work:
        def wait(self):
            return True

    def fake_all_reduce(t, op=None, group=None, async_op=False):
        calls.append((t.storage_offset(), t.numel()))
        return _Work()

    monkeypatch.setattr(dist, "all_reduce", fake_all_reduce)
    total = sum((s + 7) // 8 * 8 for s in SIZES)
    order = ["decoders", "fusion.1", "fusion.0", "tokens", "text", "vision"]        # the engine's announcement order
    for bucket_mb, expect_calls in ((0.0, None), (1e6, 2), (3 * 64 * 64 * 4 / (1 << 20), None)):
        r = GradReducer(NAMES, SIZES)
        r.world, r.bucket_bytes = 2, int(bucket_mb * (1 << 20))
        r.bind(torch.zeros(total))
        calls.clear()
        for g in order:
            r.on_ready(g)
        r.wait()
        assert not r.deferred and not r.pending
        covered = sorted(calls)
        pos = 0
        for off, n in covered:                     # disjoint cover of the whole arena
            assert off >= pos
            pos = off + n
        assert sum(n for _, n in calls) == total
        if bucket_mb == 0.0:
            assert len(calls) == sum(len(v) for v in plan_groups(NAMES, SIZES).values())
        if expect_calls is not None:
            # one huge bucket: everything up to "tokens" in one flush (merged into contiguous runs), the towers at wait()
            tower = sum((s_ + 7) // 8 * 8 for n_, s_ in zip(NAMES, SIZES) if n_.startswith("siglip_model."))
            assert len(calls) == expect_calls
            assert calls[0] == (tower, total - tower) and calls[1] == (0, tower)
        assert r.n_collectives == len(calls)


def _worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        red = GradReducer(NAMES, SIZES)
        total = sum((s + 7) // 8 * 8 for s in SIZES)
        torch.manual_seed(rank)
        arena = torch.randn(total)
        mine = arena.clone()
        red.bind(arena)
        for g in ("decoders", "fusion.1", "fusion.0", "tokens", "text", "vision"):   # backward order
            red.on_ready(g)
        red.wait()
        others = [torch.empty(total) for _ in range(world)]
        dist.all_gather(others, mine)
        assert torch.allclose(arena, sum(others), atol=1e-6)
        assert red.bytes_reduced == 4 * total
    finally:
        dist.destroy_process_group()


def test_grad_reducer_two_ranks_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port), nprocs=2, join=True)
