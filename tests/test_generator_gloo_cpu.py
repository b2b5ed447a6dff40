"""CPU, world_size-2 gloo: the multi-rank HOST logic of B200CorrespondenceGenerator - image ownership, the collective decision on
masked images, the feature all-gather, pair sharding and the result gather - with the GPU front end replaced by a deterministic
stand-in (the kernels themselves are covered by the -m gpu suite; tests/test_multigpu_gpu.py runs the real thing on two GPUs)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

K = 6


class _FakeFrontEnd:
    """detect: keypoint count and values are functions of the image content only; match: a function of the two counts."""
    device = torch.device("cpu")
    max_keypoints = K

    def _one(self, im):
        n = int(im.to(torch.int64).sum()) % K + 1
        base = float(im.to(torch.float32).mean())
        kp = torch.zeros((K, 2)); sc = torch.zeros(K); de = torch.zeros((K, 256))
        kp[:n, 0] = base + torch.arange(n); kp[:n, 1] = 2 * base; sc[:n] = base / 255.0; de[:n] = base
        return kp, sc, de, n

    def detect_pool(self, images, slots=None):
        n = max(len(images), slots or 0)
        kp = torch.zeros((n, K, 2)); sc = torch.zeros((n, K)); de = torch.zeros((n, K, 256))
        counts, shapes = [], []
        for i, im in enumerate(images):
            kp[i], sc[i], de[i], c = self._one(im)
            counts.append(c); shapes.append((int(im.shape[0]), int(im.shape[1])))
        return kp, sc, de, counts, shapes

    def detect_many(self, images):
        from gtsfm_b200.pipeline import DeviceFeatures

        kp, sc, de, counts, shapes = self.detect_pool(images)
        return [DeviceFeatures(kp[i, :c], sc[i, :c], de[i, :c], shapes[i]) for i, c in enumerate(counts)]

    def detect(self, im, mask=None):
        f = self.detect_many([im])[0]
        if mask is not None:  # stand-in for the host mask filter: drops the first keypoint
            from gtsfm_b200.pipeline import DeviceFeatures

            f = DeviceFeatures(f.kp[1:], f.score[1:], f.desc[1:], f.shape)
        return f

    def match_many(self, pairs, on_chunk=None, **kw):
        out = []
        for a, b in pairs:
            k = min(len(a), len(b))
            out.append((torch.stack([torch.arange(k), torch.arange(k)], 1).to(torch.int64), 0))
        if on_chunk:
            for c0 in range(0, len(out), 8):
                on_chunk(c0, out[c0:c0 + 8])
        return out


class _Img:
    def __init__(self, arr, mask=None):
        self.value_array, self.mask = arr, mask


def _job(with_mask):
    rng = np.random.default_rng(5)
    imgs = [_Img(rng.integers(0, 255, (8 + i, 10), dtype=np.uint8)) for i in range(7)]
    if with_mask:
        imgs[5].mask = np.ones((13, 10), np.uint8)
    graph = [(i, j) for i in range(7) for j in range(i + 1, min(7, i + 3))]  # 11 pairs; image 6 only as a second member
    return imgs, graph


def _run(with_mask):
    from gtsfm_b200.correspondence_generator import B200CorrespondenceGenerator

    gen = B200CorrespondenceGenerator(None, None, max_keypoints=K)
    gen._fe = _FakeFrontEnd()
    imgs, graph = _job(with_mask)
    kps, matches = gen.generate_correspondences(None, imgs, graph)
    return gen, kps, matches


def _worker(rank, world, port, q):
    single = {m: _run(m)[1:] for m in (False, True)}  # before the process group exists: world = 1
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = []
        for m in (False, True):
            gen, kps, matches = _run(m)
            kps1, matches1 = single[m]
            same_kp = all(np.array_equal(a.coordinates, b.coordinates) and np.array_equal(a.responses, b.responses) for a, b in zip(kps, kps1))
            same_m = sorted(matches) == sorted(matches1) and all(np.array_equal(matches[p], matches1[p]) for p in matches1)
            res.append((bool(same_kp and len(kps) == 7), bool(same_m), gen.last_detections))
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_two_gloo_ranks_equal_one_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # unmasked job: exchange path, 7 images detected once each (4 + 3); masked job: every rank detects what its pairs reference
    assert res[0][0] == (True, True, 4) and res[1][0] == (True, True, 3), res
    assert res[0][1][:2] == (True, True) and res[1][1][:2] == (True, True), res
    assert res[0][1][2] >= 4 and res[1][1][2] >= 4, res
