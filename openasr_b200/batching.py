"""Batch construction on the host side of the front-end.

The reference builds batches with ``TimeBasedSampler`` (``src/dataload/samplers.py:9-41``) and pads
them with ``load_wave_batch`` (``src/dataload/data_utils.py:126-138``): pageable ``torch.zeros`` +
per-utterance ``+=``, float32, then a synchronous ``.cuda()`` in the training loop.  Here:

* :class:`TimeBasedSampler` -- the reference's sampler, same constructor, same batches (including its
  remainder rule), so a recipe can switch the import and nothing else;
* :class:`BucketedTimeSampler` -- same duration budget, but utterances are first sorted into length
  buckets, so a batch pads to its own longest member instead of the corpus tail, and ``T`` only takes a
  few distinct values (CUDA-graph / allocator friendly);
* :class:`PinnedWaveCollator` -- pads straight into a ring of pinned host buffers (int16 PCM kept as
  int16: half the PCIe bytes, converted in kernel A) and issues the H2D copy on a side stream, so the
  copy of batch ``i+1`` overlaps the kernels of batch ``i``.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch


class TimeBasedSampler(torch.utils.data.Sampler):
    """Batches of consecutive dataset indices whose summed ``feat_length`` reaches ``duration``
    (samplers.py:9-41).  ``len(batch) % ngpu == 0`` is required to close a batch; of a trailing partial
    batch the reference keeps ``batch[b // ngpu * ngpu:]`` -- mirrored as is."""

    def __init__(self, dataset, duration=200, ngpu=1, shuffle=False):
        self.dataset = dataset
        self.dur = duration
        self.shuffle = shuffle
        self.batchs = self._build(range(len(dataset)), lambda i: dataset[i]["feat_length"], duration, ngpu)

    @staticmethod
    def _build(order: Iterable[int], length_of, budget, ngpu: int) -> List[List[int]]:
        batchs, batch, acc = [], [], 0.0
        for idx in order:
            batch.append(idx)
            acc += length_of(idx)
            if acc >= budget and len(batch) % ngpu == 0:
                batchs.append(batch)
                batch, acc = [], 0.0
        if batch:
            if len(batch) % ngpu == 0:
                batchs.append(batch)
            else:
                b = len(batch)
                batchs.append(batch[b // ngpu * ngpu:])
        return batchs

    def __iter__(self):
        if self.shuffle:
            np.random.shuffle(self.batchs)
        for b in self.batchs:
            yield b

    def __len__(self):
        return len(self.batchs)


class BucketedTimeSampler(TimeBasedSampler):
    """Length-bucketed variant: indices are sorted by ``feat_length`` (stable), cut into
    ``num_buckets`` contiguous buckets, and the reference's duration rule runs inside each bucket.
    Every index appears at most once; batches never mix buckets, so the padding a batch carries is
    bounded by the bucket's length spread."""

    def __init__(self, dataset, duration=200, ngpu=1, shuffle=False, num_buckets=8):
        self.dataset = dataset
        self.dur = duration
        self.shuffle = shuffle
        lengths = np.asarray([dataset[i]["feat_length"] for i in range(len(dataset))], dtype=np.float64)
        order = np.argsort(lengths, kind="stable")
        self.batchs = []
        for part in np.array_split(order, max(1, min(int(num_buckets), len(order)))):
            self.batchs += self._build([int(i) for i in part], lambda i: float(lengths[i]), duration, ngpu)

    def padding_fraction(self) -> float:
        """Share of padded samples over all batches (0 = no padding)."""
        lengths = [self.dataset[i]["feat_length"] for i in range(len(self.dataset))]
        used = padded = 0.0
        for b in self.batchs:
            ls = [lengths[i] for i in b]
            used += sum(ls)
            padded += max(ls) * len(ls)
        return 1.0 - used / padded if padded else 0.0


def bind_to_gpu_numa_node(device_index: int) -> Optional[List[int]]:
    """Pin the calling process to the CPUs NVML reports as local to the GPU, so that pinned staging
    buffers allocated afterwards are first-touched on the GPU's NUMA node and H2D / D2H DMA does not
    cross the socket interconnect (matters once several ranks stage batches concurrently).
    Returns the CPU list, or None when NVML / the affinity call is unavailable (no error)."""
    try:
        import os
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * i + b for i, wd in enumerate(words) for b in range(64) if (int(wd) >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None


def pad_wave_batch(waveforms: Sequence[np.ndarray], out: Optional[torch.Tensor] = None
                   ) -> Tuple[torch.Tensor, torch.Tensor]:
    """``load_wave_batch`` (data_utils.py:126-138) without the file IO: zero-pad 1-D waveforms to
    ``[B, max_len]`` + int64 lengths.  int16 input stays int16 (the kernels ingest it directly, bit-identical
    features); anything else becomes float32 like the reference.  ``out``: optional preallocated
    (pinned) ``[>= B, >= max_len]`` tensor of the right dtype to pad into."""
    lengths = [int(w.shape[0]) for w in waveforms]
    B, L = len(lengths), max(lengths)
    keep_i16 = all(w.dtype == np.int16 for w in waveforms)
    dtype = torch.int16 if keep_i16 else torch.float32
    if out is None:
        buf = torch.zeros((B, L), dtype=dtype)
    else:
        if out.dtype != dtype or out.shape[0] < B or out.shape[1] < L:
            raise ValueError("out must be a %s tensor of at least [%d, %d]" % (dtype, B, L))
        buf = out[:B, :L]
        buf.zero_()
    dst = buf.numpy()
    for i, w in enumerate(waveforms):
        dst[i, :lengths[i]] = w if keep_i16 else np.asarray(w, dtype=np.float32)
    return buf, torch.tensor(lengths, dtype=torch.int64)


class PinnedWaveCollator:
    """Ring of pinned staging buffers + a copy stream.  ``__call__(waveforms)`` pads into the next
    slot, starts the asynchronous H2D copy on the copy stream and returns
    ``(wav_cuda, lengths_cpu, ready_event)``; the consumer calls
    ``collator.wait(ready_event, wav_cuda)`` before ``SPLayer.forward``.
    A slot is reused only after its copy has completed."""

    def __init__(self, device, max_batch: int, max_len: int, slots: int = 3, int16: bool = True):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("PinnedWaveCollator needs a CUDA device (no CPU path)")
        self._dtype = torch.int16 if int16 else torch.float32
        # flat pinned slots: every batch is padded into a CONTIGUOUS [B, L] view of its slot, so the H2D copy is a
        # single asynchronous DMA (a strided [:B, :L] slice of a [max_batch, max_len] buffer would be staged through
        # a pageable temporary by torch and block the host)
        self._host = [torch.zeros((max_batch * max_len,), dtype=self._dtype).pin_memory() for _ in range(slots)]
        self._done = [None] * slots
        self._next = 0
        self._stream = torch.cuda.Stream(device=self.device)

    def __call__(self, waveforms: Sequence[np.ndarray]):
        s = self._next
        self._next = (s + 1) % len(self._host)
        if self._done[s] is not None:
            self._done[s].synchronize()
        if self._dtype == torch.float32:
            waveforms = [np.asarray(w, dtype=np.float32) for w in waveforms]
        elif not all(w.dtype == np.int16 for w in waveforms):
            raise TypeError("this collator was built for int16 PCM; pass int16=False for float input")
        B, L = len(waveforms), max(int(w.shape[0]) for w in waveforms)
        if B * L > self._host[s].numel():
            raise ValueError("batch of %d x %d samples exceeds the collator's slot (%d)" % (B, L, self._host[s].numel()))
        host, lengths = pad_wave_batch(waveforms, out=self._host[s][:B * L].view(B, L))
        assert host.is_contiguous() and host.is_pinned()
        with torch.cuda.stream(self._stream):
            dev = torch.empty(host.shape, dtype=host.dtype, device=self.device)
            dev.copy_(host, non_blocking=True)  # contiguous pinned source: one asynchronous copy
            ev = torch.cuda.Event()
            ev.record(self._stream)
        self._done[s] = ev
        return dev, lengths, ev

    def wait(self, event, wav: Optional[torch.Tensor] = None) -> None:
        """Make the current stream wait for the copy; ``wav`` (the returned tensor) is then marked as
        used on that stream so the caching allocator does not recycle it early."""
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(event)
        if wav is not None:
            wav.record_stream(cur)
