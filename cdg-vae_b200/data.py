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
    third-next batch is requested (they are views of the staging ring)."""

    def __init__(self, iterable, device, depth=3):
        self.iterable, self.device, self.depth = iterable, torch.device(device), depth

    def __len__(self):
        return len(self.iterable)

    def __iter__(self):
        dev, depth = self.device, self.depth
        side = torch.cuda.Stream(dev)
        ring = [None] * depth          # per slot: list of device buffers
        done = [None] * depth          # per slot: event after which the slot may be overwritten
        it = iter(self.iterable)
        count = 0

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
                bufs = ring[slot] = [torch.empty(s.shape, dtype=s.dtype, device=dev) for s in src]
            with torch.cuda.stream(side):
                if done[slot] is not None:
                    side.wait_event(done[slot])
                for b, s in zip(bufs, src):
                    b.copy_(s, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(side)
            return _rebuild(item, iter(bufs)), ev, slot

        nxt = load()
        prev_slot = None
        while nxt is not None:
            cur, ev, slot = nxt
            nxt = load()
            main = torch.cuda.current_stream(dev)
            main.wait_event(ev)
            yield cur
            # the consumer is back asking for the next batch: the work reading the PREVIOUS batch is enqueued
            if prev_slot is not None:
                d = torch.cuda.Event()
                d.record(torch.cuda.current_stream(dev))
                done[prev_slot] = d
            prev_slot = slot
