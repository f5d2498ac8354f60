"""Multi-GPU path (SURVEY.md row e): tiles are independent, so N ranks (one per GPU) each take a share of the tile list and rank 0
concatenates the per-tile VCF text in tile order; no collective touches the data path (only the gather of the finished text and the
max-over-ranks timing of bench.py). Here: world_size 2 over gloo on CPU with the emulation library; the gathered output must equal the
single-rank output byte for byte."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import parity_util as pu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, bam, fasta, tiles, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = [(k, t) for k, t in enumerate(tiles) if k % world == rank]
    texts = {}
    for k, t in mine:
        prev = (tiles[k - 1][0], tiles[k - 1][1], tiles[k - 1][2]) if k > 0 else (-1, 0, 0)
        out, _ = _run_one(bam, fasta, t, prev)
        texts[k] = out
    gathered = [None] * world
    dist.all_gather_object(gathered, texts)
    t = torch.tensor([float(len(mine))])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)     # the max-over-ranks reduction bench.py applies to its timings
    dist.barrier()
    if rank == 0:
        merged = {}
        for g in gathered:
            merged.update(g)
        with open(out_path, "w") as f:
            for k in range(len(tiles)):
                f.write(merged[k])
    dist.destroy_process_group()


def _run_one(bam, fasta, tile, prev):
    from uvc_b200 import capi
    bf = capi.BamFile(bam)
    rb = capi.ReadBuf()
    ctx = capi.Context(0, emulate=True)
    name, length = bf.targets[tile[0]]
    ctx.set_contig(tile[0], capi.read_fasta_contig(fasta, name))
    ctx.set_contig_name(tile[0], name)
    bf.fetch_into(rb, tile[0], max(0, tile[1] - 2000), tile[2] + 2000)
    ct = capi.make_tile(tile[0], tile[1], tile[2], tile[3], length, 0, len(rb), prev)
    ticket = ctx.submit([ct], rb.view())
    st = ctx.collect(ticket)
    text = ctx.tile_vcf(ticket, 0).decode()
    ctx.release(ticket)
    ctx.close()
    return text, st


def test_two_ranks_gloo_equal_single_rank(synth_small, tmp_path):
    tiles = [(0, 0, 3000, 4), (0, 3000, 6000, 4), (0, 6000, 9000, 4), (0, 9000, 11995, 2)]
    single, _ = pu.run_tiles(synth_small["bam"], synth_small["fasta"], tiles, True, ["vcf"])
    expect = "".join(o["vcf"] for o in single)
    out_path = str(tmp_path / "gathered.vcf")
    mp.spawn(_worker, args=(2, _free_port(), synth_small["bam"], synth_small["fasta"], tiles, out_path), nprocs=2, join=True)
    assert open(out_path).read() == expect
    assert len(expect) > 1000
