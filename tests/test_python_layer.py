"""Python layer: policy stand-in (numbering contract), channel sharding + host-side gather over a
world-size-2 gloo group on CPU, and -- on the GPU -- the Python block mirrors in a closed loop."""
import os

import numpy as np
import pytest

import lqo_py as o
import util
from liquiddsp import policy, sharding


def test_policy_numbering_matches_cognitive_engine():
    assert policy.N_CONFIGS == 616
    assert policy.config_id(0, 0, 0) == 1 and policy.config_id(10, 6, 7) == 616
    m, i, oo = policy.from_config_id(np.arange(1, 617))
    assert np.array_equal(policy.config_id(m, i, oo), np.arange(1, 617))
    p = policy.EpsilonGreedy(3, epsilon=0.0, seed=1)
    p.trials[:] = 1                                   # everything tried once with zero reward ...
    p.recompute()                                     # (the state arrays were written directly)
    p.update([0, 1, 2], [{"modulation": 8, "inner_code": 0, "outer_code": 0, "payload_valid": 1, "header_valid": 1}] * 3)
    assert p.choose() == [{"modulation": 8, "inner_code": 0, "outer_code": 0}] * 3   # ... so the rewarded one wins
    p.update([0], [{"modulation": -1, "inner_code": 0, "outer_code": 0, "payload_valid": 1, "header_valid": 1}])   # ignored


def test_channel_partition_is_exact():
    for world in (1, 2, 4, 8):
        got = np.sort(np.concatenate([sharding.channels_of_rank(4096, r, world) for r in range(world)]))
        assert np.array_equal(got, np.arange(4096))
        for r in range(world):
            ch = sharding.channels_of_rank(4096, r, world)
            assert np.array_equal(sharding.local_stream_of(ch, world), np.arange(len(ch)))


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_channels = 10
    mine = sharding.channels_of_rank(n_channels, rank, world)
    frames = []
    for s, ch in enumerate(mine):                     # two fake frames per local stream, as Rx.poll() would return them
        for seq in range(2):
            frames.append(dict(stream=s, seq=seq, sample_index=1000 * seq + ch, header_valid=1, payload_valid=seq,
                               mod_scheme=2, fec0=11, fec1=27, payload_len=1500, evm=-20.0 - ch, rssi=0.0, cfo=0.001 * ch))
    rec = sharding.records_from_frames(frames, rank, world)
    allrec = sharding.gather_records(rec, dst=0)
    if rank == 0:
        q.put(allrec)
    else:
        assert allrec is None
    dist.barrier()
    dist.destroy_process_group()


def test_host_side_gather_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    rec = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert len(rec) == 20
    assert rec["channel"].tolist() == [c for c in range(10) for _ in range(2)]
    assert rec["seq"].tolist() == [0, 1] * 10
    assert np.allclose(rec["evm"][::2], -20.0 - np.arange(10))
    assert rec["payload_valid"].sum() == 10


@pytest.mark.gpu
def test_python_blocks_closed_loop_with_policy(gpu_required):
    import liquiddsp
    rng = np.random.default_rng(6)
    tx, rx, det = liquiddsp.flex_tx(1, 0, 0), liquiddsp.flex_rx(), liquiddsp.frame_detector_cc()
    frames, infos, payloads = liquiddsp.blocks.sink(), liquiddsp.blocks.sink(), liquiddsp.blocks.sink()
    tx.msg_connect("pdus", frames, "in")
    rx.msg_connect("packet_info", infos, "in")
    rx.msg_connect("payload_data", payloads, "in")
    assert tx.message_ports_in() == ["pdus", "configuration"] and rx.message_ports_out() == ["constellation", "payload_data", "packet_info"]
    with pytest.raises(RuntimeError, match="This is not a stream block."):
        tx.work()
    pol = policy.EpsilonGreedy(1, epsilon=1.0, seed=3)
    sent = []
    for k in range(6):
        pl = rng.integers(0, 256, 256, dtype=np.uint8)
        sent.append(pl.tobytes())
        tx.post("pdus", (None, pl))
        cap = util.impair(frames.msgs[-1][1], rng, snr_db=35.0, pre=512, post=1024)
        cap = np.concatenate([cap, np.zeros((-len(cap)) % 256, np.complex64)])
        assert np.array_equal(det.work(cap), cap)
        rx.work(cap)
        pol.update([0], [infos.msgs[-1]])
        tx.post("configuration", pol.choose()[0])            # cognitive_engine -> flex_tx contract
    assert [p[1] for p in payloads.msgs] == sent
    assert all(i["header_valid"] == 1 and i["payload_valid"] == 1 for i in infos.msgs)
    assert len({(i["modulation"], i["inner_code"], i["outer_code"]) for i in infos.msgs}) > 1   # the loop really reconfigured
    assert det.num_frames >= 6


def test_capture_file_helpers(tmp_path):
    # raw complex64 capture files (GNU Radio file_sink format), chunking in multiples of 256, PDU view of a frame
    from liquiddsp import replay
    x = (np.arange(3000) + 1j * np.arange(3000)).astype(np.complex64)
    path = tmp_path / "cap.c32"
    x.tofile(path)
    one = replay.open_capture(str(path))
    assert one.shape == (1, 3000) and np.array_equal(one[0], x)
    two = replay.open_capture(str(path), n_channels=2)
    assert two.shape == (2, 1500) and np.array_equal(two[1], x[1500:])
    assert replay.chunk_bounds(1000, 512) == [(0, 512), (512, 1000)]
    with pytest.raises(ValueError):
        replay.chunk_bounds(1000, 100)
    fr = {"header_valid": 1, "payload_valid": 1, "payload": b"abc", "mod_scheme": 27, "fec0": 11, "fec1": 27,
          "framesyms": np.ones(4, np.complex64)}
    msgs = replay.to_pdus(fr)
    assert [m[0] for m in msgs] == ["constellation", "payload_data", "packet_info"]
    assert msgs[2][1] == {"header_valid": 1, "payload_valid": 1, "modulation": 8, "inner_code": 1, "outer_code": 2}
    assert [m[0] for m in replay.to_pdus(dict(fr, header_valid=0))] == ["constellation"]


@pytest.mark.gpu
def test_capture_replay_in_chunks_equals_one_shot(gpu_required, tmp_path):
    # two capture files of different lengths replayed in 256-multiples through the pipelined receiver:
    # same frames, same order per stream, as one call over the whole captures
    import lqo_py as o
    import util
    from liquiddsp import capi, replay
    rng = np.random.default_rng(7)
    caps = []
    for s_ in range(2):
        frames = [o.tx_frame(util.PSK4, util.CRC24, 11, 27, rng.integers(0, 256, 200 + 100 * k, dtype=np.uint8)) for k in range(3 + s_)]
        caps.append(util.build_capture(frames, rng, [900] * len(frames), snr_db=20.0, cfo=0.01, tau=0.1))
        caps[-1].tofile(tmp_path / ("ch%d.c32" % s_))
    files = [replay.open_capture(str(tmp_path / ("ch%d.c32" % s_)))[0] for s_ in range(2)]
    got = list(replay.replay(files, chunk=4096))
    rx = capi.Rx(2)
    rx.execute(caps)
    ref = rx.poll()
    key = lambda f: (f["stream"], f["sample_index"])
    assert sorted(map(key, got)) == sorted(map(key, ref)) and len(ref) == 7
    for s_ in range(2):
        a = [f for f in got if f["stream"] == s_]
        b = [f for f in ref if f["stream"] == s_]
        assert [f["payload"] for f in a] == [f["payload"] for f in b] and all(f["payload_valid"] for f in a)
        assert [f["sample_index"] for f in a] == sorted(f["sample_index"] for f in a)


@pytest.mark.gpu
def test_single_capture_replay_sharded_equals_chunked_replay(gpu_required, tmp_path):
    # one capture file (one channel): the time-sharded replay returns the frames of the chunk-by-chunk replay
    import lqo_py as o
    import util
    from liquiddsp import replay
    rng = np.random.default_rng(17)
    frames = [o.tx_frame(util.QAM16, util.CRC24, 11, 1, rng.integers(0, 256, 150 + 60 * (k % 5), dtype=np.uint8)) for k in range(25)]
    cap = util.build_capture(frames, rng, [600 + 350 * (k % 4) for k in range(25)], snr_db=22.0, cfo=-0.008, tau=0.2)
    cap.tofile(tmp_path / "one.c32")
    x = replay.open_capture(str(tmp_path / "one.c32"))[0]
    ref = list(replay.replay([x], chunk=2048))
    got = replay.replay_sharded(x, workers=16, seg_len=8192, preroll=4096)
    assert len(ref) == 25 and all(f["payload_valid"] for f in ref)
    assert [(f["sample_index"], f["payload"]) for f in got] == [(f["sample_index"], f["payload"]) for f in ref]
    assert [m[0] for f in got for m in replay.to_pdus(f)] == ["constellation", "payload_data", "packet_info"] * 25


def test_pdu_tagged_stream_adapters_round_trip():
    """PDUs -> tagged stream -> PDUs: lengths and contents survive any chunking, tags carry
    absolute offsets, the chunker only produces multiples of 256."""
    from liquiddsp import adapters
    rng = np.random.default_rng(3)
    pdus = [(rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64) for n in (2690, 1, 700, 28282)]
    src = adapters.pdu_to_tagged_stream()
    for p_ in pdus:
        src.post("pdus", (None, p_))
    assert src.pending() == sum(map(len, pdus))
    sink = adapters.tagged_stream_to_pdu()
    got, all_tags = [], []
    while src.pending():
        items, tags = src.work(int(rng.integers(1, 5000)))
        all_tags += tags
        got += sink.work(items, tags)
    assert [t[0] for t in all_tags] == [0, 2690, 2691, 3391] and all(t[1] == "packet_len" for t in all_tags)
    assert [t[2] for t in all_tags] == [len(p_) for p_ in pdus]
    assert len(got) == 4 and all(np.array_equal(g[1], p_) for g, p_ in zip(got, pdus))
    src.post("pdus", (None, pdus[0]))
    chunks = list(adapters.stream_chunks(src, chunk=1024, idle_gap=600))
    assert all(len(c) == 1024 for c in chunks) and len(chunks) == 3 + 1
    assert np.array_equal(np.concatenate(chunks)[:2690], pdus[0]) and not np.concatenate(chunks)[2690:].any()
    with pytest.raises(ValueError):
        list(adapters.stream_chunks(src, chunk=100))


@pytest.mark.gpu
def test_flex_tx_pdus_through_the_adapters_into_flex_rx(gpu_required):
    """The authors' loopback, wired without GNU Radio: flex_tx PDUs -> pdu_to_tagged_stream -> 256-multiples -> flex_rx."""
    import liquiddsp
    from liquiddsp import adapters
    from liquiddsp.blocks import sink
    rng = np.random.default_rng(4)
    tx, rx = liquiddsp.flex_tx(1, 1, 2), liquiddsp.flex_rx()
    stream, payloads, infos = adapters.pdu_to_tagged_stream(), sink(), sink()
    tx.msg_connect("pdus", stream, "pdus")
    rx.msg_connect("payload_data", payloads, "in")
    rx.msg_connect("packet_info", infos, "in")
    sent = [rng.integers(0, 256, 300, dtype=np.uint8) for _ in range(4)]
    for p_ in sent:
        tx.post("pdus", (None, p_))
        stream.post("pdus", (None, np.zeros(700, np.complex64)))          # idle time between frames
    for chunk in adapters.stream_chunks(stream, chunk=4096, idle_gap=4096):
        rx.work(chunk)
    assert [m[1] for m in payloads.msgs] == [p_.tobytes() for p_ in sent]
    assert all(i["payload_valid"] == 1 and i["modulation"] == 1 and i["inner_code"] == 1 and i["outer_code"] == 2 for i in infos.msgs)
