"""Micro-batcher in front of `CaptionPipeline`.

The reference serves one video per request and serialises requests with a semaphore of one
(server/services/task_manager.py:22, inference_service.py:59-60), so its GPU sees batches of one.  Here single-video
requests from any number of threads are collected for at most `max_delay_ms`, stacked per frame shape into pinned host
batches of up to `max_batch` videos, and fed to the three-stream pipeline; every request gets a future that resolves to
its own token ids.  A video's result does not depend on the batch it rides in (rows are independent everywhere on the
path; tests check it), so batching is invisible to the caller.  The REST layer of the reference stays where it is: its
handler calls `batcher.submit(frames).result()` instead of `engine.infer(...)`.

All CUDA work is issued from the batcher's one worker thread.
"""
from __future__ import annotations

import queue
import threading
import time
from concurrent.futures import Future
from typing import Dict, List, Tuple

import torch


def group_requests(shapes: List[Tuple[int, ...]], max_batch: int) -> List[List[int]]:
    """Indices of requests that can share a batch: same [T,H,W,3] shape, arrival order kept, at most max_batch per group."""
    groups: Dict[Tuple[int, ...], List[List[int]]] = {}
    order: List[List[int]] = []
    for i, shp in enumerate(shapes):
        lst = groups.setdefault(tuple(shp), [])
        if not lst or len(lst[-1]) >= max_batch:
            lst.append([])
            order.append(lst[-1])
        lst[-1].append(i)
    return order


class _PinnedRing:
    """Pinned staging for the batches of ONE frame shape [T,H,W,3]: `depth` buffers of `max_batch` videos each, used in rotation
    (a buffer is not rewritten while its H2D copy may still be running) and sliced to the size of the batch at hand."""

    def __init__(self, frame_shape: Tuple[int, ...], max_batch: int, depth: int):
        self.bufs = [torch.empty((max_batch,) + tuple(frame_shape), dtype=torch.uint8).pin_memory() for _ in range(depth)]
        self.i = 0
        self.bytes = sum(b.numel() for b in self.bufs)

    def next(self, n: int) -> torch.Tensor:
        b = self.bufs[self.i % len(self.bufs)]
        self.i += 1
        return b[:n]


class MicroBatcher:
    def __init__(self, model, max_batch: int = 64, max_delay_ms: float = 5.0, max_new_tokens: int = 20, decode_group: int = 2,
                 max_pinned_bytes: int = 4 << 30):
        self.model = model
        self.max_batch, self.max_delay = int(max_batch), float(max_delay_ms) / 1e3
        self.max_new = int(max_new_tokens)
        self._q: "queue.Queue" = queue.Queue()
        self._closed = False
        self._decode_group = int(decode_group)
        self._max_pinned = int(max_pinned_bytes)  # page-locked staging memory is bounded: least recently used frame shapes are dropped
        self.batches = 0                      # statistics: batches submitted / requests served
        self.served = 0
        self.error: BaseException | None = None   # set if the worker died; every pending and later request fails with it
        self._thread = threading.Thread(target=self._run, name="vcb200-microbatcher", daemon=True)
        self._thread.start()

    # ------------------------------------------------------------------ client side
    def submit(self, frames_u8: torch.Tensor) -> Future:
        """frames_u8: uint8 [T,H,W,3] host tensor of ONE video (any frame size; resized on the GPU like the reference's
        transform).  Returns a Future whose result is (ids: list[int] without the eos padding, length)."""
        if self.error is not None:
            raise RuntimeError("MicroBatcher worker failed") from self.error
        if self._closed:
            raise RuntimeError("MicroBatcher is closed")
        if frames_u8.dtype != torch.uint8 or frames_u8.ndim != 4 or frames_u8.shape[-1] != 3:
            raise ValueError(f"expect uint8 [T,H,W,3], got {frames_u8.dtype} {tuple(frames_u8.shape)}")
        fut: Future = Future()
        self._q.put((frames_u8, fut))
        if self.error is not None and not fut.done():          # the worker died between the check above and the put
            fut.set_exception(self.error)
        return fut

    def close(self) -> None:
        self._closed = True
        self._q.put(None)
        self._thread.join()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ------------------------------------------------------------------ worker
    def _gather(self, block: bool):
        """Requests that arrived within max_delay of the first one (or nothing, when not blocking and the queue is empty)."""
        reqs = []
        try:
            first = self._q.get(block=block, timeout=0.05 if block else None)
        except queue.Empty:
            return reqs, False
        if first is None:
            return reqs, True
        reqs.append(first)
        deadline = time.monotonic() + self.max_delay
        while len(reqs) < 4 * self.max_batch:
            left = deadline - time.monotonic()
            if left <= 0:
                break
            try:
                r = self._q.get(timeout=left)
            except queue.Empty:
                break
            if r is None:
                return reqs, True
            reqs.append(r)
        return reqs, False

    def _run(self) -> None:
        """Worker thread.  Any exception that escapes the serving loop (a failed pinned allocation, a CUDA error in the pipeline)
        fails every queued and in-flight request and closes the batcher — a dead worker must never leave a Future hanging."""
        inflight: list = []                   # (ticket, futures)
        try:
            self._serve(inflight)
        except BaseException as e:            # noqa: BLE001
            self.error = e
            self._closed = True
            for _, futs in inflight:
                for f in futs:
                    if not f.done():
                        f.set_exception(e)
            while True:
                try:
                    r = self._q.get_nowait()
                except queue.Empty:
                    break
                if r is not None and not r[1].done():
                    r[1].set_exception(e)

    def _serve(self, inflight: list) -> None:
        pipe = self.model.pipeline(max_new_tokens=self.max_new, decode_group=self._decode_group)
        from collections import OrderedDict
        rings: "OrderedDict[Tuple[int, ...], _PinnedRing]" = OrderedDict()     # keyed by the FRAME shape, least recently used first
        stop = False

        def ring_for(frame_shape: Tuple[int, ...]) -> _PinnedRing:
            r = rings.pop(frame_shape, None)
            if r is None:
                r = _PinnedRing(frame_shape, self.max_batch, pipe.depth + 1)
            rings[frame_shape] = r                                              # most recently used last
            while len(rings) > 1 and sum(x.bytes for x in rings.values()) > self._max_pinned:
                # the oldest shape's buffers may still feed a copy in flight: finish what is in flight before dropping them
                resolve(len(inflight))
                rings.popitem(last=False)
            return r

        def resolve(n: int) -> None:
            for _ in range(min(n, len(inflight))):
                ticket, futs = inflight.pop(0)
                try:
                    ids, lens = pipe.result(ticket)
                    for b, f in enumerate(futs):
                        n_tok = int(lens[b])
                        f.set_result((ids[b, :n_tok].tolist(), n_tok))
                    self.served += len(futs)
                except Exception as e:                       # noqa: BLE001 — hand the failure to every waiter
                    for f in futs:
                        if not f.done():
                            f.set_exception(e)

        while not stop:
            reqs, stop = self._gather(block=not inflight)
            if reqs:
                for group in group_requests([tuple(r[0].shape) for r in reqs], self.max_batch):
                    buf = ring_for(tuple(reqs[group[0]][0].shape)).next(len(group))
                    for j, i in enumerate(group):
                        buf[j].copy_(reqs[i][0])
                    futs = [reqs[i][1] for i in group]
                    try:
                        if len(inflight) >= pipe.depth - 1:
                            resolve(1)
                        inflight.append((pipe.submit(buf, to_host=True), futs))
                        self.batches += 1
                    except Exception as e:                   # noqa: BLE001
                        for f in futs:
                            f.set_exception(e)
            if inflight and (self._q.empty() or stop):
                resolve(len(inflight))                       # idle: nobody else to wait for, finish what is in flight
        resolve(len(inflight))
        pipe.drain()
