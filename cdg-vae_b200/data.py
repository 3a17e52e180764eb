"""Input staging for the train loops: copy batch i+1 host->device on a side stream while batch i
computes (the reference moves each batch with a synchronous pageable `.cuda()`, modules/train.py:163-165)."""
import torch


def _to(item, device):
    if torch.is_tensor(item):
        return item.to(device, non_blocking=True)
    if isinstance(item, (tuple, list)):
        return type(item)(_to(t, device) for t in item)
    return item


def _record(item, stream):
    if torch.is_tensor(item):
        item.record_stream(stream)
    elif isinstance(item, (tuple, list)):
        for t in item:
            _record(t, stream)


class DevicePrefetcher:
    """Wrap any iterable of (nested tuples of) host tensors; yields the same structure on `device`.
    Pinned host tensors make the copies truly asynchronous."""

    def __init__(self, iterable, device):
        self.iterable, self.device = iterable, torch.device(device)

    def __len__(self):
        return len(self.iterable)

    def __iter__(self):
        side = torch.cuda.Stream(self.device)
        it = iter(self.iterable)

        def load():
            try:
                item = next(it)
            except StopIteration:
                return None
            with torch.cuda.stream(side):
                out = _to(item, self.device)
                ev = torch.cuda.Event()
                ev.record(side)
            return out, ev

        nxt = load()
        while nxt is not None:
            cur, ev = nxt
            nxt = load()
            main = torch.cuda.current_stream(self.device)
            main.wait_event(ev)
            _record(cur, main)
            yield cur
