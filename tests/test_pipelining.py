"""Several tickets outstanding on one context (the schedule bench.py's end-to-end phase and a pipelined host use): batch k+1 is submitted before
batch k is collected, scored and turned into text. The VCF text of every tile must equal what the same tiles give when each batch is run to
completion before the next one starts (which test_vcf_parity.py pins against the reference)."""
import pytest

import parity_util as pu
from uvc_b200 import capi


def _tiles_text(info, tile_groups, emulate, pipelined, register=False):
    bf = capi.BamFile(info["bam"])
    rb = capi.ReadBuf()
    ctx = capi.Context(0, emulate=emulate)
    for tid, (name, _) in enumerate(bf.targets):
        ctx.set_contig(tid, capi.read_fasta_contig(info["fasta"], name))
        ctx.set_contig_name(tid, name)
    groups, prev = [], (-1, 0, 0)
    for g in tile_groups:
        ct = []
        for (tid, beg, end, flag) in g:
            r0 = len(rb)
            bf.fetch_into(rb, tid, max(0, beg - 2000), end + 2000)
            ct.append(capi.make_tile(tid, beg, end, flag, bf.targets[tid][1], r0, len(rb), prev))
            prev = (tid, beg, end)
        groups.append(ct)
    view = rb.view()
    if register:        # page-locked caller buffers: the library uploads the records straight from them (no staging copy)
        import ctypes as C
        ctx.lib.uvcgpu_host_register_reads.argtypes = [C.c_void_p]
        assert ctx.lib.uvcgpu_host_register_reads(C.byref(view)) == 0

    def finish(ticket, n):
        st = ctx.collect(ticket)
        ctx.score(ticket)
        text = [ctx.tile_vcf(ticket, k).decode() for k in range(n)]
        assert bytes(ctx.batch_vcf(ticket)).decode() == "".join(text)      # the batch-level call is the concatenation in tile order
        ctx.release(ticket)
        return text, st

    out, launches = [], 0
    if pipelined:
        tickets = [ctx.submit(g, view) for g in groups]        # all batches enqueued before the first one is collected
        for t, g in zip(tickets, groups):
            text, st = finish(t, len(g))
            out += text
            launches += int(st.gpu_launches)
    else:
        for g in groups:
            text, st = finish(ctx.submit(g, view), len(g))
            out += text
            launches += int(st.gpu_launches)
    if register:
        ctx.lib.uvcgpu_host_unregister_reads.argtypes = [C.c_void_p]
        ctx.lib.uvcgpu_host_unregister_reads(C.byref(view))
    ctx.close()
    rb.close()
    bf.close()
    return out, launches


GROUPS = [[(0, 0, 3000, 0), (0, 3000, 6000, 0)], [(0, 6000, 9000, 0)], [(0, 9000, 11995, 0)]]


def test_emulation_tickets_in_flight(synth_small):
    a, _ = _tiles_text(synth_small, GROUPS, True, False)
    b, _ = _tiles_text(synth_small, GROUPS, True, True)
    assert a == b and any(len(t) > 0 for t in a)


@pytest.mark.gpu
def test_cuda_tickets_in_flight(synth_small):
    a, la = _tiles_text(synth_small, GROUPS, False, False)
    b, lb = _tiles_text(synth_small, GROUPS, False, True)
    assert la > 0 and lb > 0
    assert a == b and any(len(t) > 0 for t in a)


@pytest.mark.gpu
def test_cuda_registered_host_buffers(synth_small, synth_umi):
    """Records uploaded straight from page-locked caller buffers (several tiles per batch share one source; the offsets are rebased on the device)."""
    a, _ = _tiles_text(synth_small, GROUPS, False, True)
    b, lb = _tiles_text(synth_small, GROUPS, False, True, register=True)
    assert lb > 0 and a == b and any(len(t) > 0 for t in a)
    g = [[(0, 1000, 2500, 0)], [(0, 2500, 4000, 0)]]
    assert _tiles_text(synth_umi, g, False, False)[0] == _tiles_text(synth_umi, g, False, True, register=True)[0]
