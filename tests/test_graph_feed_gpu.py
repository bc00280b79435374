"""Captured chains (trainner_redux_b200/chain_graph.py) and the side-stream prefetcher: a replayed CUDA graph must
produce exactly what the eager path produces for the same Philox position, fresh noise on every replay, the crop the
host drew for that step — and the prefetcher's static slots must hand feed_data the right bytes."""

import pytest
import torch

from oracle import otf_oracle as O
from trainner_redux_b200 import degradations as D
from trainner_redux_b200.prefetch import CUDAPrefetcher
from trainner_redux_b200.realesrgan_feed import OTFOptions, RealESRGANFeed

pytestmark = pytest.mark.gpu


def _opt(noise: str, b: int, **kw) -> OTFOptions:
    g = 1.0 if noise == "gaussian" else 0.0
    base = dict(scale=4, gt_size=64, blur_prob=1, blur_prob2=1, gaussian_noise_prob=g, noise_range=(1, 30), poisson_scale_range=(0.05, 3),
                gray_noise_prob=0.4, gaussian_noise_prob2=g, noise_range2=(1, 25), poisson_scale_range2=(0.05, 2.5), gray_noise_prob2=0.4,
                jpeg_range=(30, 95), jpeg_range2=(30, 95), resize_prob=(0, 0, 1), resize_mode_list=["bicubic"], resize_mode_prob=[1],
                resize_prob2=(0, 0, 1), resize_mode_list2=["bilinear"], resize_mode_prob2=[1], resize_mode_list3=["area"],
                resize_mode_prob3=[1], final_jpeg_first_prob=0.0, queue_size=4 * b)
    base.update(kw)
    return OTFOptions(**base)


def _data(b: int, size: int, seed: int, dev=None) -> dict:
    d = {"gt": O.synth_gt(b, size, size, "natural", seed=seed), "kernel1": O.synth_blur_kernels(b, seed=seed),
         "kernel2": O.synth_blur_kernels(b, seed=seed + 50), "sinc_kernel": O.synth_sinc_or_pulse(b, seed=seed)}
    return {k: v.to(dev) for k, v in d.items()} if dev is not None else d


@pytest.mark.parametrize("noise", ["gaussian", "poisson"])
@pytest.mark.parametrize("use_pool", [False, True])
def test_replayed_graph_equals_eager_and_draws_fresh_noise(noise, use_pool, dev):
    b = 4
    data = _data(b, 96, 3, dev)  # device-resident: the addresses repeat, so the plan shape is captured on its 2nd sighting
    outs = {}
    for graphs in (True, False):
        feed = RealESRGANFeed(_opt(noise, b), device=dev, manual_seed=11, use_pool=use_pool)
        feed.use_graphs = graphs
        steps = []
        for _ in range(7):
            feed.feed_data(data)
            steps.append((feed.gt.clone(), feed.lq.clone(), dict(feed.last_plan)))
        outs[graphs] = steps
        if graphs:
            assert feed.graphs.captures == 1 and feed.graphs.hits >= 5, (feed.graphs.captures, feed.graphs.hits)
        else:
            assert feed.graphs.captures == 0
    for i, ((g0, l0, p0), (g1, l1, p1)) in enumerate(zip(outs[True], outs[False])):
        assert p0["crop"] == p1["crop"]
        assert torch.equal(g0, g1), f"step {i}: GT crop (graph vs eager)"
        assert torch.equal(l0, l1), f"step {i}: LQ (graph vs eager) for the same Philox position"
    if not use_pool:  # every replay draws fresh noise and follows the crop drawn for that step
        lqs = [l for _, l, _ in outs[True]]
        assert all(not torch.equal(lqs[i], lqs[i + 1]) for i in range(len(lqs) - 1))
        assert len({p["crop"] for _, _, p in outs[True]}) > 1


def test_same_counter_same_field_different_counter_different_field(dev):
    """The device-side offset word: a captured noise launch follows the word, not the value at capture time."""
    x = torch.rand(2, 3, 32, 32, device=dev)
    sig = torch.full((2,), 20.0, device=dev)
    word = torch.zeros(1, dtype=torch.int64, device=dev)
    out = torch.empty_like(x)
    from trainner_redux_b200 import _lib

    def launch():
        _lib.call("otf_gaussian_noise_f32", _lib.ptr(x), 2, 3, 32, 32, _lib.ptr(sig), None, None, None, 1234, 5, _lib.ptr(word), 1,
                  _lib.ptr(out), _lib.stream())

    launch()  # warm-up outside the capture
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        launch()
    got = []
    for w in (0, 1, 1, 7):
        word.fill_(w)
        g.replay()
        got.append(out.clone())
    assert torch.equal(got[1], got[2]) and not torch.equal(got[0], got[1]) and not torch.equal(got[1], got[3])
    gen = D.PhiloxState(seed=1234, offset=5 + 7)  # eager with the host-side offset 5 + 7 == graph with word 7
    assert torch.equal(D.add_gaussian_noise_pt(x, sig, 0, True, False, generator=gen), got[3])


def test_noise_field_injection_is_the_reference_tail(dev):
    xc = torch.rand(3, 3, 20, 28) * 1.2 - 0.1  # expectations on the CPU: ATen's CUDA `x / 255.0` multiplies by a reciprocal
    fc = torch.randn(3, 3, 20, 28) * 0.05
    x, f = xc.to(dev), fc.to(dev)
    assert torch.equal(D.add_noise_field_pt(x, f).cpu(), (xc + fc).clamp(0, 1))
    assert torch.equal(D.add_noise_field_pt(x, f, clip=True, rounds=True).cpu(), ((xc + fc) * 255.0).round().clamp(0, 255) / 255.0)
    assert torch.equal(D.add_noise_field_pt(x, f, clip=False).cpu(), xc + fc)


def test_prefetcher_static_slots_and_feed(dev):
    b, n = 4, 7
    batches = [_data(b, 96, 100 + i) for i in range(n)]
    for d in batches:
        d["gt"] = d["gt"].pin_memory()
    direct = RealESRGANFeed(_opt("gaussian", b), device=dev, manual_seed=5, use_pool=False)
    direct.use_graphs = False
    want = []
    for d in batches:
        direct.feed_data(d)
        want.append((direct.gt.clone(), direct.lq.clone()))
    feed = RealESRGANFeed(_opt("gaussian", b), device=dev, manual_seed=5, use_pool=False)
    pf = CUDAPrefetcher(batches, device=dev, slots=2)
    got, ptrs = [], set()
    batch = pf.next()
    while batch is not None:
        assert batch["gt"].is_cuda and pf.h2d_bytes == sum(v.numel() * 4 for v in batch.values())
        ptrs.add(batch["gt"].data_ptr())
        feed.feed_data(batch)
        got.append((feed.gt.clone(), feed.lq.clone()))
        batch = pf.next()
    assert len(got) == n and len(ptrs) == 2, "two static slots, reused round-robin"
    assert feed.graphs.captures == 2 and feed.graphs.hits >= 1  # one captured chain per slot
    for i, ((g0, l0), (g1, l1)) in enumerate(zip(got, want)):
        assert torch.equal(g0, g1) and torch.equal(l0, l1), f"batch {i}"
    pf.reset()
    assert torch.equal(pf.next()["gt"].cpu(), batches[0]["gt"])


def test_uint8_gt_through_the_prefetcher(dev):
    b = 4
    d = _data(b, 96, 9)
    d8 = dict(d, gt=(d["gt"] * 255.0).round().clamp(0, 255).to(torch.uint8))
    pf = CUDAPrefetcher([d8, d8, d8], device=dev)
    feed = RealESRGANFeed(_opt("gaussian", b), device=dev, manual_seed=5, use_pool=False)
    ref = RealESRGANFeed(_opt("gaussian", b), device=dev, manual_seed=5, use_pool=False)
    ref.use_graphs = False
    batch = pf.next()
    while batch is not None:
        assert batch["gt"].dtype == torch.uint8
        feed.feed_data(batch)
        ref.feed_data(dict(d, gt=d8["gt"].float() / 255.0))
        assert torch.equal(feed.lq, ref.lq) and torch.equal(feed.gt, ref.gt)
        batch = pf.next()


def test_readback_side_stream(dev):
    """CUDAReadback: the values that land in the pinned ring are the tensor's at the time of the call, even when the
    producer overwrites the device buffer two calls later (static outputs of a captured chain)."""
    from trainner_redux_b200.prefetch import CUDAReadback

    rb = CUDAReadback(dev, depth=2)
    buf = [torch.empty(3, 64, 64, device=dev) for _ in range(2)]
    hosts = []
    for i in range(6):
        buf[i % 2].fill_(float(i))
        h = rb.read(buf[i % 2])
        assert h.is_pinned() and h.shape == buf[0].shape
        hosts.append((i, h))
        if i >= 1:  # the ring is two deep: the previous result is still intact after this call
            rb.wait()
            assert torch.all(hosts[-1][1] == float(i)) and torch.all(hosts[-2][1] == float(i - 1))
    with pytest.raises(RuntimeError):
        rb.read(torch.zeros(4))


def test_readback_as_bytes_is_lossless_on_the_lattice(dev):
    """CUDAReadback.read(as_u8=True): an image on the 8-bit lattice (every level, any length incl. a ragged tail) crosses
    as ``clamp(round(x * 255), 0, 255)`` bytes (`tensor2img`'s conversion, img_util.py:112-181, done on the device) and
    ``u8.float() / 255`` restores the fp32 tensor bit for bit; off-lattice and out-of-range values round half to even and
    saturate like the reference's ``(img * 255.0).round()`` + uint8 clamp; a finished LQ batch survives the trip."""
    from trainner_redux_b200 import _lib as L
    from trainner_redux_b200.prefetch import CUDAReadback

    rb = CUDAReadback(dev, depth=2)
    for n in (256, 1027, 64 * 3 * 56 * 56 + 3):
        lv = (torch.arange(n) * 7919) % 256
        x = (lv.float() / 255.0).to(dev)
        h = rb.read(x, as_u8=True)
        rb.wait()
        assert h.dtype == torch.uint8 and h.is_pinned() and torch.equal(h, lv.to(torch.uint8))
        assert torch.equal(h.float() / 255.0, x.cpu())
    z = (torch.arange(1030) % 256).float().div(255.0).to(dev)[3:]  # a view that does not start on a 16-byte boundary
    h = rb.read(z, as_u8=True)
    rb.wait()
    assert torch.equal(h.float() / 255.0, z.cpu())
    y = torch.tensor([-0.3, 0.0, 0.5 / 255, 1.5 / 255, 2.5 / 255, 0.49999, 1.0, 1.7, 254.5 / 255], device=dev)
    u = torch.empty(y.numel(), dtype=torch.uint8, device=dev)
    L.call("otf_f32_to_u8", L.ptr(y), y.numel(), L.ptr(u), L.stream())
    want = torch.clamp(torch.round(y.cpu() * 255.0), 0, 255).to(torch.uint8)
    assert torch.equal(u.cpu(), want)
    feed = RealESRGANFeed(_opt("gaussian", 4), device=dev, manual_seed=5, use_pool=False)
    data = _data(4, 96, 3, dev)
    for _ in range(3):  # eager, capture, replay: the read-back follows the producer on the current stream each time
        feed.feed_data(data)
        h = rb.read(feed.lq, as_u8=True)
        rb.wait()
        assert torch.equal(h.float() / 255.0, feed.lq.cpu())
    with pytest.raises(TypeError):
        rb.read(torch.zeros(8, dtype=torch.float64, device=dev), as_u8=True)


def test_gt_window_is_a_view_like_the_reference(dev):
    """paired_random_crop returns the GT window as a slice of the batch (transforms.py:124-129): feed_data hands out the
    same view (no bytes move) unless the pool / MoA / ``gt_view = False`` ask for a dense copy — identical values."""
    b = 4
    data = _data(b, 96, 3, dev)
    got = {}
    for name, kw in (("view", {}), ("copy", {"gt_view": False}), ("pool", {"use_pool": True})):
        feed = RealESRGANFeed(_opt("gaussian", b), device=dev, manual_seed=11, use_pool=kw.get("use_pool", False))
        feed.gt_view = kw.get("gt_view", True)
        steps = []
        for _ in range(5):  # eager, eager, then replays of the captured chain
            feed.feed_data(data)
            top, left = feed.last_plan["crop"]
            if name == "view":
                assert not feed.gt.is_contiguous() and feed.gt.untyped_storage().data_ptr() == data["gt"].untyped_storage().data_ptr()
                assert feed.gt.storage_offset() == 4 * top * 96 + 4 * left and feed.gt.shape == (b, 3, 64, 64)
            else:
                assert feed.gt.is_contiguous() and feed.gt.untyped_storage().data_ptr() != data["gt"].untyped_storage().data_ptr()
            assert feed.lq.is_contiguous()
            steps.append((feed.gt.clone(), feed.lq.clone()))
        got[name] = steps
    for (g0, l0), (g1, l1) in zip(got["view"], got["copy"]):
        assert torch.equal(g0, g1) and torch.equal(l0, l1)
    # (the pool hands out its own batches; while it fills, the pair passes through)
    assert torch.equal(got["pool"][0][0], got["view"][0][0])
