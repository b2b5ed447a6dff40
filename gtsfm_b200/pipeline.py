"""Device-resident batched front-end: detect -> top-k -> describe -> match -> verify with every tensor kept in HBM.

This is the L2 seam of SURVEY.md §8b (`CorrespondenceGeneratorBase.generate_correspondences`): instead of one Dask task
per image and per pair, each pickling ~5 MB of features (det_desc_correspondence_generator.py:65-85), one process per
GPU walks its shard of the visibility graph and only small results (match index arrays, E / R / t) leave the device.
PyTorch is used for device buffers and the stream only; all arithmetic is libgtsfm_b200.so through the `*_dev` C ABI.
"""
from __future__ import annotations

import contextlib
import os
import threading
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, weights
from .detector_descriptor import KEYPOINT_THRESHOLD, NMS_RADIUS, REMOVE_BORDERS, SuperPointEngine
from .verifier import DEFAULT_SEED, E_MAX_ITERS, RANSAC_SUCCESS_PROB


@dataclass
class DeviceFeatures:
    kp: torch.Tensor  # (k, 2) float32 (x, y), device
    score: torch.Tensor  # (k,)
    desc: torch.Tensor  # (k, 256)
    shape: Tuple[int, int]

    def __len__(self) -> int:
        return int(self.kp.shape[0])


# concurrent LightGlue instances used by match_many.  Default 1: measured on B200 (40-pair steps of 5000 x 5000 keypoints) 215.8 /
# 224.5 / 217.0 pairs/s with 1 / 2 / 3 lanes - the matcher's persistent kernels already fill the GPU, unlike SuperPoint's
MATCH_LANES = int(os.environ.get("B2_MATCH_LANES", "1"))
# concurrent SuperGlue instances used by match_superglue_many: 117.9 / 122.7 / 129.3 / 130.9 pairs/s with 1 / 2 / 3 / 4 lanes (B200,
# 40-pair steps at 5000 keypoints, RANSAC on its own stream)
SG_LANES = int(os.environ.get("B2_SG_LANES", "3"))
DETECT_LANES = int(os.environ.get("B2_DETECT_LANES", "4"))  # concurrent SuperPoint instances used by detect_many
RESERVE_SMS_FOR_VERIFY = int(os.environ.get("B2_RESERVE_SMS", "8"))  # k_rs_hyp_E keeps 16 x 64-thread CTAs busy for ~1 ms


class DeviceFrontEnd:
    def __init__(self, superpoint_sd, lightglue_sd=None, device: int = 0, max_keypoints: int = 5000, cpu_semantics: bool = True,
                 ctx: Optional[_lib.Context] = None, superglue_sd=None, fp16_attention: bool = False):
        if not torch.cuda.is_available():
            raise _lib.B200Error("DeviceFrontEnd needs a CUDA device; there is no CPU fallback")
        self.device = torch.device("cuda", device)
        self.ctx = ctx or _lib.Context(device)
        self.lib = self.ctx.lib
        self.max_keypoints = max_keypoints
        self.prune_min = -1 if cpu_semantics else 1536
        self.fp16_attention = 1 if fp16_attention else 0  # opt-in: the reference's CUDA numerics (lightglue.py:116-121)
        blob = weights.pack_superpoint(weights.load_state_dict(superpoint_sd))
        self._sp_blob = blob
        self._lanes = []  # extra (context, stream) SuperPoint lanes of detect_many
        self._counts = None
        self.ctx.check(self.lib.b2_superpoint_set_weights(self.ctx.handle, _lib.ptr(blob), blob.size), "superpoint_set_weights")
        self._lg_blob = None
        self._mlanes = []  # extra (context, stream, executor) LightGlue lanes of match_many
        self._mpool0 = None
        self._reserve_sms = 0
        if lightglue_sd is not None:
            blob = weights.pack_lightglue(weights.load_state_dict(lightglue_sd))
            self._lg_blob = blob
            self.ctx.check(self.lib.b2_lightglue_set_weights(self.ctx.handle, _lib.ptr(blob), blob.size), "lightglue_set_weights")
        self._sg_blob = None
        self._sglanes = []  # (context, stream, executor) SuperGlue lanes of match_superglue_many (lane 0 = self.ctx)
        if superglue_sd is not None:
            blob = weights.pack_superglue(weights.load_state_dict(superglue_sd))
            self._sg_blob = blob
            self.ctx.check(self.lib.b2_superglue_set_weights(self.ctx.handle, _lib.ptr(blob), blob.size), "superglue_set_weights")
        # verification runs on its own context + stream + host thread so that the (latency-bound, 16-CTA) RANSAC kernels of
        # pair p overlap the matcher kernels of pair p+1 (ctypes calls release the GIL)
        self._vctx: Optional[_lib.Context] = None
        self._vstream: Optional[torch.cuda.Stream] = None
        self._vpool = None
        self._vlock = threading.Lock()

    # measurement helpers over every context that runs SuperPoint / matcher kernels for this front end (bench.py)
    def _all_ctx(self):
        return [self.ctx] + [c for c, _ in self._lanes] + [l[0] for l in self._mlanes] + [l[0] for l in self._sglanes[1:]]

    def launch_count(self) -> int:
        return sum(c.launch_count() for c in self._all_ctx())

    def profile_start(self, kernel_prefix: str) -> None:
        for c in self._all_ctx():
            c.profile_start(kernel_prefix)

    def profile_stop(self):
        ms, n, w = 0.0, 0, 0.0
        for c in self._all_ctx():
            a, b, d = c.profile_stop()
            ms, n, w = ms + a, n + b, w + d
        return ms, n, w

    def _stream(self):
        return _lib.C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def ingest(self, image: torch.Tensor, max_resolution: int = 760) -> torch.Tensor:
        """Loader-side down-size on the device (gtsfm/loader/loader_base.py:160-200 -> gtsfm/utils/images.py:102-129,150-220):
        uint8 (H, W[, C]) device tensor -> cubic resize so that the short side is `max_resolution` (unchanged if already smaller)."""
        assert image.dtype == torch.uint8 and image.is_cuda and image.is_contiguous()
        h, w = int(image.shape[0]), int(image.shape[1])
        ch = 1 if image.dim() == 2 else int(image.shape[2])
        if min(h, w) <= max_resolution:
            return image
        if h <= w:
            nh, nw = max_resolution, int(np.round(w * (max_resolution / float(h))).astype(np.int32))
        else:
            nh, nw = int(np.round(h * (max_resolution / float(w))).astype(np.int32)), max_resolution
        out = torch.empty((nh, nw) if image.dim() == 2 else (nh, nw, ch), dtype=torch.uint8, device=self.device)
        rc = self.lib.b2_image_resize_dev(self.ctx.handle, _lib.ptr(image), h, w, ch, w * ch, _lib.ptr(out), nh, nw, self._stream())
        self.ctx.check(rc, "image_resize_dev")
        return out

    def detect(self, image: torch.Tensor, mask: Optional[np.ndarray] = None) -> DeviceFeatures:
        """image: uint8 device tensor (H, W) or (H, W, 3|4), contiguous.  One C call (detect -> device top-k -> describe): no
        torch kernels on the path, and the dense map never outlives the call (interleaving images on one handle is safe).
        `mask` (host (H, W) array, 1 = keep; gtsfm/common/keypoints.py:112-127) is applied BEFORE the top-k like the
        reference wrapper does (gtsfm/.../superpoint.py:87-91); masked images take the two-call path."""
        assert image.dtype == torch.uint8 and image.is_cuda and image.is_contiguous()
        h, w = int(image.shape[0]), int(image.shape[1])
        ch = 1 if image.dim() == 2 else int(image.shape[2])
        if mask is not None:
            return self._detect_masked(image, h, w, ch, np.asarray(mask))
        k = self.max_keypoints
        kp = torch.empty((k, 2), dtype=torch.float32, device=self.device)
        score = torch.empty(k, dtype=torch.float32, device=self.device)
        desc = torch.empty((k, 256), dtype=torch.float32, device=self.device)
        n = _lib.C.c_int(0)
        rc = self.lib.b2_superpoint_extract_dev(self.ctx.handle, _lib.ptr(image), h, w, ch, w * ch, KEYPOINT_THRESHOLD, NMS_RADIUS,
                                                REMOVE_BORDERS, k, _lib.ptr(kp), _lib.ptr(score), _lib.ptr(desc), _lib.C.byref(n),
                                                self._stream())
        self.ctx.check(rc, "superpoint_extract_dev")
        return DeviceFeatures(kp[: n.value], score[: n.value], desc[: n.value], (h, w))

    def detect_many(self, images) -> list:
        """`detect` for a list of images with every image enqueued before the first result is read: no host synchronisation
        between images (b2_superpoint_extract_async_dev), one at the end (b2_superpoint_finish_dev).  Images alternate between
        DETECT_LANES library contexts on as many streams, so one image's narrow kernels (80-CTA layers, the one-block top-k / scan) and
        launch gaps are filled by the other's.  Same outputs as [detect(im) for im in images]."""
        if len(images) == 0:
            return []
        kp_all, score_all, desc_all, counts, shapes = self.detect_pool(images)
        return [DeviceFeatures(kp_all[i, :n], score_all[i, :n], desc_all[i, :n], shapes[i]) for i, n in enumerate(counts)]

    def detect_pool(self, images, slots: Optional[int] = None):
        """detect_many's engine: -> (kp (slots, k, 2), score (slots, k), desc (slots, k, 256) device tensors, counts list, shapes);
        image i fills slot i (`slots` >= len(images): padding slots for the multi-GPU feature exchange)."""
        k = self.max_keypoints
        if self._counts is None or self._counts.numel() < len(images):  # page-locked once, reused (cudaHostAlloc is slow)
            self._counts = torch.zeros(max(64, len(images)), dtype=torch.int32).pin_memory()
        counts = self._counts
        outs = []
        lanes = [(self.ctx, torch.cuda.current_stream(self.device))]
        for j in range(1, min(DETECT_LANES, max(1, len(images)))):
            if len(self._lanes) < j:  # another SuperPoint instance (own work buffers) + its stream
                ctx = _lib.Context(self.device.index or 0)
                ctx.check(self.lib.b2_superpoint_set_weights(ctx.handle, _lib.ptr(self._sp_blob), self._sp_blob.size), "superpoint_set_weights")
                self._lanes.append((ctx, torch.cuda.Stream(self.device)))
            self._lanes[j - 1][1].wait_stream(lanes[0][1])  # the images were produced on the caller's stream
            lanes.append(self._lanes[j - 1])
        # ONE allocation per output kind for the whole list (a fresh 5 MB descriptor block per image is a cudaMalloc each when the
        # caller keeps every image's features alive: 360 of them cost 0.7 s for 120 frames), sliced per image
        nimg = max(len(images), slots or 0)
        kp_all = torch.empty((nimg, k, 2), dtype=torch.float32, device=self.device)
        score_all = torch.empty((nimg, k), dtype=torch.float32, device=self.device)
        desc_all = torch.empty((nimg, k, 256), dtype=torch.float32, device=self.device)
        for _, stream in lanes[1:]:
            stream.wait_stream(lanes[0][1])  # the pool's previous owner (caching allocator) finished on the caller's stream
            for t in (kp_all, score_all, desc_all):
                t.record_stream(stream)
        for i, image in enumerate(images):
            assert image.dtype == torch.uint8 and image.is_cuda and image.is_contiguous()
            ctx, stream = lanes[i % len(lanes)]
            h, w = int(image.shape[0]), int(image.shape[1])
            ch = 1 if image.dim() == 2 else int(image.shape[2])
            kp, score, desc = kp_all[i], score_all[i], desc_all[i]
            rc = self.lib.b2_superpoint_extract_async_dev(ctx.handle, _lib.ptr(image), h, w, ch, w * ch, KEYPOINT_THRESHOLD,
                                                          NMS_RADIUS, REMOVE_BORDERS, k, _lib.ptr(kp), _lib.ptr(score), _lib.ptr(desc),
                                                          _lib.C.c_void_p(counts.data_ptr() + 4 * i), _lib.C.c_void_p(stream.cuda_stream))
            ctx.check(rc, "superpoint_extract_async_dev")
            outs.append((h, w))
        for ctx, stream in lanes:
            ctx.check(self.lib.b2_superpoint_finish_dev(ctx.handle, _lib.C.c_void_p(stream.cuda_stream)), "superpoint_finish_dev")
        return kp_all, score_all, desc_all, counts[: len(images)].tolist(), outs

    def _detect_masked(self, image: torch.Tensor, h: int, w: int, ch: int, mask: np.ndarray) -> DeviceFeatures:
        cap = SuperPointEngine.capacity(h, w)
        xy = torch.empty((cap, 2), dtype=torch.float32, device=self.device)
        sc = torch.empty(cap, dtype=torch.float32, device=self.device)
        n, tok = _lib.C.c_int(0), _lib.C.c_uint64(0)
        rc = self.lib.b2_superpoint_detect_dev(self.ctx.handle, _lib.ptr(image), h, w, ch, w * ch, KEYPOINT_THRESHOLD, NMS_RADIUS,
                                               REMOVE_BORDERS, _lib.ptr(xy), _lib.ptr(sc), cap, _lib.C.byref(n), _lib.C.byref(tok),
                                               self._stream())
        self.ctx.check(rc, "superpoint_detect_dev")
        nk = min(n.value, cap)
        hxy, hsc = xy[:nk].cpu().numpy(), sc[:nk].cpu().numpy()
        r = np.round(hxy).astype(int)
        keep = np.flatnonzero(mask[r[:, 1], r[:, 0]] == 1)
        if len(keep) > self.max_keypoints:  # the k largest responses, ties by lower index, row-major order kept
            keep = np.sort(keep[np.argsort(-hsc[keep], kind="stable")[: self.max_keypoints]])
        kp = torch.from_numpy(np.ascontiguousarray(hxy[keep])).to(self.device)
        score = torch.from_numpy(np.ascontiguousarray(hsc[keep])).to(self.device)
        desc = torch.empty((len(keep), 256), dtype=torch.float32, device=self.device)
        rc = self.lib.b2_superpoint_describe_dev(self.ctx.handle, tok.value, _lib.ptr(kp), len(keep), _lib.ptr(desc), self._stream())
        self.ctx.check(rc, "superpoint_describe_dev")
        return DeviceFeatures(kp, score, desc, (h, w))

    def match(self, a: DeviceFeatures, b: DeviceFeatures, depth_confidence=0.95, width_confidence=0.99, filter_threshold=0.1):
        cap = max(1, min(len(a), len(b)))
        out = torch.empty((cap, 2), dtype=torch.int64, device=self.device)
        k, stop = _lib.C.c_int(0), _lib.C.c_int(0)
        prm = _lib.LightGlueParams(depth_confidence, width_confidence, filter_threshold, self.prune_min, self.fp16_attention)
        rc = self.lib.b2_lightglue_match_dev(self.ctx.handle, _lib.ptr(a.kp), _lib.ptr(a.desc), len(a), _lib.ptr(b.kp), _lib.ptr(b.desc),
                                             len(b), _lib.C.byref(prm), _lib.ptr(out), None, _lib.C.byref(k), _lib.C.byref(stop),
                                             self._stream())
        self.ctx.check(rc, "lightglue_match_dev")
        return out[: k.value], stop.value

    def match_superglue_many(self, pairs: Sequence[Tuple[DeviceFeatures, DeviceFeatures]], on_pair=None, **kw) -> List[torch.Tensor]:
        """`match_superglue` over a list of pairs, dealt round-robin to SG_LANES library contexts (own SuperGlue instance, stream
        and host thread each): SuperGlue runs pair by pair with many narrow kernels (2000-keypoint GNN layers, one-block
        filters), which concurrent pairs fill.  `on_pair(index, matches)` is called from the lane's thread as a pair completes."""
        from concurrent.futures import ThreadPoolExecutor

        lanes = min(SG_LANES, len(pairs))
        if lanes <= 1:
            out = []
            for i, (a, b) in enumerate(pairs):
                m = self.match_superglue(a, b, **kw)
                if on_pair:
                    on_pair(i, m)
                out.append(m)
            return out
        while len(self._sglanes) < lanes:  # lane 0 included: every lane has its own host thread
            if not self._sglanes:
                ctx = self.ctx
            else:
                ctx = _lib.Context(self.device.index or 0)
                ctx.check(self.lib.b2_superglue_set_weights(ctx.handle, _lib.ptr(self._sg_blob), self._sg_blob.size), "superglue_set_weights")
                ctx.set_option("reserve_sms", self._reserve_sms)
            self._sglanes.append((ctx, torch.cuda.Stream(self.device) if self._sglanes else None, ThreadPoolExecutor(max_workers=1)))
        main = torch.cuda.current_stream(self.device)

        def work(lane, idxs):
            ctx, stream, _ = self._sglanes[lane]
            res = []
            with torch.cuda.stream(stream if stream is not None else main):
                for i in idxs:
                    m = self.match_superglue(*pairs[i], ctx=ctx, **kw)
                    if stream is not None:
                        m.record_stream(main)
                    if on_pair:
                        on_pair(i, m)
                    res.append((i, m))
            return res

        for _, stream, _ in self._sglanes[1:lanes]:
            stream.wait_stream(main)  # the features were produced on the caller's stream
        futs = [self._sglanes[l][2].submit(work, l, list(range(l, len(pairs), lanes))) for l in range(lanes)]
        out: List[Optional[torch.Tensor]] = [None] * len(pairs)
        for f in futs:
            for i, m in f.result():
                out[i] = m
        return out  # type: ignore[return-value]

    def match_superglue(self, a: DeviceFeatures, b: DeviceFeatures, sinkhorn_iters: int = 20, match_threshold: float = 0.2,
                        ctx: Optional[_lib.Context] = None) -> torch.Tensor:
        """SuperGlue on device-resident features -> (k, 2) int64 device tensor (rows (i, matches0[i]) ascending in i)."""
        cap = max(1, min(len(a), len(b)))
        out = torch.empty((cap, 2), dtype=torch.int32, device=self.device)  # the ABI writes uint32 rows
        k = _lib.C.c_int(0)
        ctx = ctx or self.ctx
        rc = self.lib.b2_superglue_match_dev(ctx.handle, _lib.ptr(a.kp), _lib.ptr(a.score), _lib.ptr(a.desc), len(a), a.shape[0], a.shape[1],
                                             _lib.ptr(b.kp), _lib.ptr(b.score), _lib.ptr(b.desc), len(b), b.shape[0], b.shape[1],
                                             int(sinkhorn_iters), float(match_threshold), _lib.ptr(out), None, _lib.C.byref(k), self._stream())
        ctx.check(rc, "superglue_match_dev")
        return out[: k.value].to(torch.int64)

    def match_many(self, pairs: Sequence[Tuple[DeviceFeatures, DeviceFeatures]], on_chunk=None, **kw) -> List[Tuple[torch.Tensor, int]]:
        """`match_batch` over any number of pairs: lock-step batches of 8, dealt to MATCH_LANES library contexts (own LightGlue
        instance, stream and host thread each) so that one batch's per-layer host syncs, small glue kernels and launch gaps
        are filled by another batch's kernels.  `on_chunk(first_pair_index, results)` is called (from the lane's thread) as soon
        as a batch is complete - the hook bench.py uses to start verification early.  Results in input order."""
        from concurrent.futures import ThreadPoolExecutor

        chunks = [(c0, pairs[c0:c0 + 8]) for c0 in range(0, len(pairs), 8)]
        lanes = min(MATCH_LANES, len(chunks))
        if lanes <= 1:
            out = []
            for c0, ch in chunks:
                r = self.match_batch(ch, **kw)
                if on_chunk:
                    on_chunk(c0, r)
                out += r
            return out
        while len(self._mlanes) < lanes - 1:  # extra lanes: context + LightGlue weights + stream + a one-thread executor
            ctx = _lib.Context(self.device.index or 0)
            ctx.check(self.lib.b2_lightglue_set_weights(ctx.handle, _lib.ptr(self._lg_blob), self._lg_blob.size), "lightglue_set_weights")
            ctx.set_option("reserve_sms", self._reserve_sms)
            self._mlanes.append((ctx, torch.cuda.Stream(self.device), ThreadPoolExecutor(max_workers=1)))
        if self._mpool0 is None:
            self._mpool0 = ThreadPoolExecutor(max_workers=1)
        main = torch.cuda.current_stream(self.device)

        def work(lane, c0, ch):
            if lane == 0:
                with torch.cuda.stream(main):
                    r = self.match_batch(ch, **kw)
            else:
                ctx, stream, _ = self._mlanes[lane - 1]
                with torch.cuda.stream(stream):
                    r = self.match_batch(ch, ctx=ctx, **kw)
                for m, _ in r:
                    m.record_stream(main)
            if on_chunk:
                on_chunk(c0, r)
            return r

        for _, stream, _ in self._mlanes[: lanes - 1]:
            stream.wait_stream(main)  # the features were produced on the caller's stream
        futs = []
        for i, (c0, ch) in enumerate(chunks):
            lane = i % lanes
            ex = self._mpool0 if lane == 0 else self._mlanes[lane - 1][2]
            futs.append(ex.submit(work, lane, c0, ch))
        out = []
        for f in futs:
            out += f.result()
        return out

    def match_batch(self, pairs: Sequence[Tuple[DeviceFeatures, DeviceFeatures]], depth_confidence=0.95, width_confidence=0.99,
                    filter_threshold=0.1, ctx: Optional[_lib.Context] = None) -> List[Tuple[torch.Tensor, int]]:
        """LightGlue over a list of pairs through `b2_lightglue_match_batched_dev`: the library walks up to 8 pairs in
        lock-step (one launch per layer step for all their images).  -> [(matches (k, 2) int64 device tensor, stop layer)]."""
        n = len(pairs)
        if n == 0:
            return []
        arr = (_lib.LightGluePair * n)()
        outs = []
        for i, (a, b) in enumerate(pairs):
            out = torch.empty((max(1, min(len(a), len(b))), 2), dtype=torch.int64, device=self.device)
            outs.append(out)
            arr[i].kp0, arr[i].desc0, arr[i].n0 = a.kp.data_ptr(), a.desc.data_ptr(), len(a)
            arr[i].kp1, arr[i].desc1, arr[i].n1 = b.kp.data_ptr(), b.desc.data_ptr(), len(b)
            arr[i].out_matches, arr[i].out_scores = out.data_ptr(), None
        prm = _lib.LightGlueParams(depth_confidence, width_confidence, filter_threshold, self.prune_min, self.fp16_attention)
        ctx = ctx or self.ctx
        rc = self.lib.b2_lightglue_match_batched_dev(ctx.handle, arr, n, _lib.C.byref(prm), self._stream())
        ctx.check(rc, "lightglue_match_batched_dev")
        return [(outs[i][: arr[i].out_k], int(arr[i].out_stop_layer)) for i in range(n)]

    def verify_async(self, a: DeviceFeatures, b: DeviceFeatures, matches: torch.Tensor, cal1, cal2, threshold_px: float = 4.0,
                     seed: int = DEFAULT_SEED):
        """Same as verify() but returns a concurrent.futures.Future; `matches` must already be complete on the device
        (match() synchronises its stream before returning)."""
        from concurrent.futures import ThreadPoolExecutor

        with self._vlock:
            self._ensure_verify_lane()
        return self._vpool.submit(self.verify, a, b, matches, cal1, cal2, threshold_px, seed, self._vctx, self._vstream)

    def _ensure_verify_lane(self):
        from concurrent.futures import ThreadPoolExecutor

        if self._vpool is None:
            self._vctx = _lib.Context(self.device.index)
            self._vstream = torch.cuda.Stream(self.device)
            self._vpool = ThreadPoolExecutor(max_workers=1)
            # the matcher's persistent kernels (one CTA per SM) leave a few SMs to the concurrent RANSAC kernels: a CTA
            # that finds its SM occupied would wait for a whole CTA lifetime and double the kernel's duration
            self._reserve_sms = RESERVE_SMS_FOR_VERIFY
            for c in [self.ctx] + [l[0] for l in self._mlanes] + [l[0] for l in self._sglanes[1:]]:
                c.set_option("reserve_sms", RESERVE_SMS_FOR_VERIFY)

    def verify(self, a: DeviceFeatures, b: DeviceFeatures, matches: torch.Tensor, cal1: Sequence[float], cal2: Sequence[float],
               threshold_px: float = 4.0, seed: int = DEFAULT_SEED, ctx: Optional[_lib.Context] = None,
               stream: Optional[torch.cuda.Stream] = None):
        """cal = (f, u0, v0).  -> (E (3,3) | None, R, t, num_inliers, mask device tensor)."""
        ctx = ctx or self.ctx
        sptr = _lib.C.c_void_p(stream.cuda_stream) if stream is not None else self._stream()
        k = int(matches.shape[0])
        matches = matches.contiguous()  # kept alive in this frame until the C call has returned
        with torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext():
            mask = torch.zeros(max(k, 1), dtype=torch.uint8, device=self.device)  # zero-filled on the stream the kernels use
        if k < 6:  # opencv_verifier_base.py:70-79
            return None, None, None, 0, mask[:k]
        E, R, t = np.zeros(9), np.zeros(9), np.zeros(3)
        c1, c2 = np.asarray(cal1, np.float64), np.asarray(cal2, np.float64)
        n = _lib.C.c_int(0)
        prm = _lib.RansacParams(threshold_px / max(c1[0], c2[0]), RANSAC_SUCCESS_PROB, E_MAX_ITERS, seed)
        rc = self.lib.b2_ransac_essential_dev(ctx.handle, _lib.ptr(a.kp), _lib.ptr(b.kp), _lib.ptr(matches), k, _lib.ptr(c1),
                                              _lib.ptr(c2), _lib.C.byref(prm), _lib.ptr(E), _lib.ptr(mask), _lib.C.byref(n), _lib.ptr(R),
                                              _lib.ptr(t), sptr)
        ctx.check(rc, "ransac_essential_dev")
        if rc == 1:
            return None, None, None, 0, mask[:k]
        return E.reshape(3, 3), R.reshape(3, 3), t, n.value, mask[:k]


from .distributed import shard_pairs  # noqa: E402,F401  (re-export: `p mod world` partitioner)
