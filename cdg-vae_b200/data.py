"""Input staging for the train loops: copy batch i+1 host->device on a side stream while batch i
computes (the reference moves each batch with a synchronous pageable `.cuda()`, modules/train.py:163-165).

Three fixed sets of device buffers rotate, so the steady state allocates nothing and the copy engine runs
back to back.  A consumer may look one batch ahead (modules/train.py::_lookahead does, to know which batch
is the last), so when it asks for batch i+1 it is only guaranteed to have enqueued the work of batch i-1:
that is the slot whose "may be overwritten" event is recorded then; with three slots the copy of batch i+1
still never waits in steady state."""
import torch


def _leaves(item, out):
    if torch.is_tensor(item):
        out.append(item)
    elif isinstance(item, (tuple, list)):
        for t in item:
            _leaves(t, out)
    return out


def _rebuild(item, it):
    if torch.is_tensor(item):
        return next(it)
    if isinstance(item, (tuple, list)):
        return type(item)(_rebuild(t, it) for t in item)
    return item


class DevicePrefetcher:
    """Wrap any iterable of (nested tuples of) host tensors; yields the same structure on `device`.
    Pinned host tensors make the copies truly asynchronous.  The yielded tensors are only valid until the
    third-next batch is requested (they are views of the staging ring).

    pixels=True: uint8 tensors are image bytes in the dataset's native format (what modules/datasets.py:24-27 reads
    from the PNGs).  They cross PCIe as bytes (1 instead of 4 per pixel) and are converted on the device, on the copy
    stream, by `cdg_pixels_to_float` — the same `(p - 127.5) / 127.5` in float64 rounded to fp32 that the reference's
    dataset applies on the host (datasets.py:28, :42) — so the consumer sees bit-identical fp32 batches."""

    def __init__(self, iterable, device, depth=3, pixels=False, convert_on=None):
        import os
        self.iterable, self.device, self.depth, self.pixels = iterable, torch.device(device), depth, pixels
        # where the byte -> fp32 kernel runs: "copy" (behind its DMA on the copy stream; default), "own" (a high-priority
        # stream of its own) or "compute" (on the consumer's stream, right before the step).  Measured on B200 in a 12-step
        # loop of the bench workload (tools/e2e_diag.py, copy alone 40.6 ms, step alone 28.0 ms): 38.8 / 48.3 / 49.1 ms per
        # step -- the other two keep the byte staging buffer busy until the consumer's step has run, which delays the next DMA
        self.convert_on = convert_on or os.environ.get("CDG_PREFETCH_CONVERT", "copy")
        # len() is exact when the wrapped loader's is (lists, DeviceDataLoader): lets the train loops know the last batch
        # without reading one batch ahead of the step they enqueue
        self.exact_len = isinstance(iterable, (list, tuple)) or getattr(iterable, "exact_len", False)

    def __len__(self):
        return len(self.iterable)

    def __iter__(self):
        dev, depth = self.device, self.depth
        side = torch.cuda.Stream(dev)
        ring = [None] * depth          # per slot: list of device buffers
        done = [None] * depth          # per slot: event after which the slot may be overwritten
        decoded = [None] * depth       # per slot: fp32 buffers of the uint8 leaves (pixels=True), else None
        it = iter(self.iterable)
        count = 0
        conv = side if self.convert_on == "copy" else torch.cuda.Stream(dev, priority=-1)

        def convert(pairs, stream):
            from . import _lib
            for b, d in pairs:
                _lib.check(_lib.lib().cdg_pixels_to_float(b.data_ptr(), b.numel(), d.data_ptr(), stream.cuda_stream))

        def load():
            nonlocal count
            try:
                item = next(it)
            except StopIteration:
                return None
            slot = count % depth
            count += 1
            src = _leaves(item, [])
            bufs = ring[slot]
            if bufs is None or len(bufs) != len(src) or any(b.shape != s.shape or b.dtype != s.dtype for b, s in zip(bufs, src)):
                # allocate ON the copy stream: the caching allocator hands out blocks assuming they are next touched on the
                # stream that is current at allocation time.  Allocated on the compute stream, a ring buffer could reuse
                # the block of a temporary that the still-running previous step had just freed (its noise tensor, say), and
                # the copy stream would overwrite it under that step (seen as a 1e-4 drift in step 1's loss).
                with torch.cuda.stream(side):
                    bufs = ring[slot] = [torch.empty(s.shape, dtype=s.dtype, device=dev) for s in src]
                    decoded[slot] = [torch.empty(s.shape, dtype=torch.float32, device=dev)
                                     if (self.pixels and s.dtype == torch.uint8) else None for s in src]
            outs = [b if d is None else d for b, d in zip(bufs, decoded[slot])]
            with torch.cuda.stream(side):
                if done[slot] is not None:
                    side.wait_event(done[slot])
                for b, s in zip(bufs, src):
                    b.copy_(s, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(side)
            pending = [(b, d) for b, d in zip(bufs, decoded[slot]) if d is not None]
            if pending and self.convert_on != "compute":
                with torch.cuda.stream(conv):
                    conv.wait_event(ev)
                    convert(pending, conv)
                    ev = torch.cuda.Event()
                    ev.record(conv)
                pending = []
            return _rebuild(item, iter(outs)), ev, slot, pending

        nxt = load()
        prev_slot = None
        while nxt is not None:
            cur, ev, slot, pending = nxt
            nxt = load()
            main = torch.cuda.current_stream(dev)
            main.wait_event(ev)
            if pending:
                convert(pending, main)
            for t in _leaves(cur, []):
                t.record_stream(main)      # read on the compute stream: the allocator must not recycle it under that work
            yield cur
            # the consumer is back asking for the next batch: the work reading the PREVIOUS batch is enqueued
            if prev_slot is not None:
                d = torch.cuda.Event()
                d.record(torch.cuda.current_stream(dev))
                done[prev_slot] = d
            prev_slot = slot


class DeviceDataLoader:
    """Device-resident replacement for `DataLoader(dataset, batch_size, shuffle=True)` over the reference's
    map-style datasets (modules/datasets.py:14-65: `x_data` [, `y_data`] numpy arrays converted per item with
    `torch.FloatTensor`).  The whole dataset is converted to fp32 and uploaded ONCE; every epoch yields the same
    batches, in the same order, as the reference's DataLoader would under the same global RNG state:

      * `iter()` draws the loader's base seed (torch/utils/data/dataloader.py, _BaseDataLoaderIter.__init__),
      * the first `next()` draws the sampler seed and a `torch.randperm(n)` from it (RandomSampler.__iter__),

    so noise draws that follow (`torch.randn` in model.encode, modules/model.py:276) see an identical RNG stream.
    Batches are gathered on the device: no per-item conversion, no collate, no host->device copy per step.

    `pixels=True`: the FIRST array is the dataset's raw uint8 image bytes (what the reference's datasets hold before
    modules/datasets.py:28); it stays uint8 in HBM (4x more images per GB than the fp32 copy) and every batch is assembled by
    one kernel that gathers the sampler's rows and applies the identical `(p - 127.5) / 127.5` (cdg_pixels_gather_to_float):
    the batches are bit-identical to the fp32 route's.
    """

    def __init__(self, *arrays, batch_size, shuffle=True, drop_last=False, device="cuda", unpack_single=True, pixels=False,
                 prefetch=False):
        self.pixels = bool(pixels)
        self.prefetch = bool(prefetch)
        if self.pixels:
            first = torch.as_tensor(arrays[0])
            if first.dtype != torch.uint8:
                raise TypeError("pixels=True expects the dataset's uint8 image bytes")
            self.tensors = [first.contiguous().to(device)] + [torch.as_tensor(a).to(dtype=torch.float32).to(device) for a in arrays[1:]]
            self._row_bytes = int(first[0].numel()) if first.shape[0] else 0
            if self._row_bytes % 16 != 0:
                raise ValueError("pixels=True needs images of a whole number of 16-byte units")
        else:
            self.tensors = [torch.as_tensor(a).to(dtype=torch.float32).to(device) for a in arrays]
        n = self.tensors[0].shape[0]
        assert all(t.shape[0] == n for t in self.tensors)
        self.n, self.batch_size, self.shuffle, self.drop_last = n, int(batch_size), shuffle, drop_last
        self.unpack_single = unpack_single
        self.device = torch.device(device)
        self.exact_len = True

    @classmethod
    def from_dataset(cls, dataset, batch_size, shuffle=True, drop_last=False, device="cuda"):
        """`dataset`: a reference LabeledDataset / UnLabeledDataset / TabularDataset (has x_data and maybe y_data)."""
        arrays = [dataset.x_data] + ([dataset.y_data] if hasattr(dataset, "y_data") else [])
        return cls(*arrays, batch_size=batch_size, shuffle=shuffle, drop_last=drop_last, device=device)

    def _gather_pixels(self, idx):
        import ctypes as C
        from . import _lib
        img = self.tensors[0]
        idx = idx.to(dtype=torch.int64).contiguous()
        out = torch.empty((idx.numel(),) + tuple(img.shape[1:]), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().cdg_pixels_gather_to_float(C.c_void_p(img.data_ptr()), img.shape[0], self._row_bytes,
                                                             C.c_void_p(idx.data_ptr()), idx.numel(), C.c_void_p(out.data_ptr()),
                                                             C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        return out

    def __len__(self):
        return self.n // self.batch_size if self.drop_last else (self.n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        torch.empty((), dtype=torch.int64).random_()                     # the DataLoader iterator's base seed
        return self._batches()

    def _batches(self):
        if self.shuffle:
            seed = int(torch.empty((), dtype=torch.int64).random_().item())
            g = torch.Generator()
            g.manual_seed(seed)
            perm = torch.randperm(self.n, generator=g).to(self.device, non_blocking=True)
        else:
            perm = None
        ranges = []
        for lo in range(0, self.n, self.batch_size):
            hi = min(self.n, lo + self.batch_size)
            if self.drop_last and hi - lo < self.batch_size:
                break
            ranges.append((lo, hi))

        def assemble(lo, hi):
            if self.pixels:
                idx = perm[lo:hi] if perm is not None else torch.arange(lo, hi, device=self.device)
                items = [self._gather_pixels(idx)] + [t.index_select(0, idx) for t in self.tensors[1:]]
            elif perm is None:
                items = [t[lo:hi] for t in self.tensors]
            else:
                idx = perm[lo:hi]
                items = [t.index_select(0, idx) for t in self.tensors]
            return items

        def pack(items):
            return items[0] if (len(items) == 1 and self.unpack_single) else tuple(items)

        if not (self.pixels and self.device.type == "cuda" and self.prefetch):
            for lo, hi in ranges:
                yield pack(assemble(lo, hi))
            return
        # prefetch=True: batch i + 1 is gathered + converted on a side stream while the consumer's work on batch i runs.  OFF by
        # default: measured on B200 at 131,072-image batches it is SLOWER than assembling in front of the step (25.1 vs 24.0 ms
        # per step: the 10 GB gather pass takes SMs and HBM bandwidth away from the step's persistent GEMM kernels); it pays
        # when the consumer is not GPU-bound
        side = self.__dict__.get("_side")
        if side is None:
            side = self.__dict__["_side"] = torch.cuda.Stream(device=self.device)

        def launch(lo, hi):
            side.wait_stream(torch.cuda.current_stream(self.device))     # (the permutation's upload; the previous epoch's state)
            with torch.cuda.stream(side):
                items = assemble(lo, hi)
                ev = torch.cuda.Event()
                ev.record(side)
            return items, ev

        nxt = launch(*ranges[0]) if ranges else None
        for i in range(len(ranges)):
            items, ev = nxt
            nxt = launch(*ranges[i + 1]) if i + 1 < len(ranges) else None
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            for t in items:
                t.record_stream(cur)              # allocated on the side stream, read on the consumer's
            yield pack(items)
