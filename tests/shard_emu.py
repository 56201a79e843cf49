"""Lock-step emulation of `world` ranks inside one process (one thread per rank, one GPU): every rank runs up to a
collective, the collective is computed once, every rank continues.  Lets the sharded arithmetic be compared with a single
index on one device; the NCCL plumbing itself is exercised by `bench.py --gpus N` and the gloo tests."""
import threading

import torch


class Emulator:
    def __init__(self, world: int):
        self.world = world
        self.barrier = threading.Barrier(world)
        self.slots = {}
        self.lock = threading.Lock()
        # one rank runs at a time (the turn is handed over inside the collectives): the ranks share one GPU, one stream
        # and therefore the per-stream scratch buffers of aura_snn_rag_b200.ops, which must not be used concurrently
        self.turn = threading.Lock()

    def collectives(self, rank: int):
        state = {"n": 0}
        world, slots, lock, turn = self.world, self.slots, self.lock, self.turn

        class _Barrier:                      # wait without holding the turn
            @staticmethod
            def wait():
                turn.release()
                try:
                    self.barrier.wait()
                finally:
                    turn.acquire()
        barrier = _Barrier

        def all_reduce(t):
            key = ("r", state["n"]); state["n"] += 1
            with lock:
                slots.setdefault(key, []).append(t)
            barrier.wait()
            if rank == 0:
                total = torch.stack([x.double() if x.is_floating_point() else x for x in slots[key]]).sum(0)
                for x in slots[key]:
                    x.copy_(total.to(x.dtype))
            barrier.wait()
            return t

        def all_gather(t):
            key = ("g", state["n"]); state["n"] += 1
            with lock:
                slots.setdefault(key, {})[rank] = t
            barrier.wait()
            out = torch.stack([slots[key][r] for r in range(world)])
            barrier.wait()
            return out
        return all_reduce, all_gather

    def run(self, fn):
        """fn(rank, all_reduce, all_gather) on `world` threads; returns the list of results, re-raises a failure."""
        results, errors = [None] * self.world, []

        def body(rank):
            self.turn.acquire()
            try:
                torch.cuda.set_device(0)
                ar, ag = self.collectives(rank)
                results[rank] = fn(rank, ar, ag)
            except BaseException as e:          # noqa: BLE001 - surface it in the main thread
                errors.append(e)
                self.barrier.abort()
            finally:
                self.turn.release()
        threads = [threading.Thread(target=body, args=(r,)) for r in range(self.world)]
        [t.start() for t in threads]
        [t.join() for t in threads]
        torch.cuda.synchronize()
        if errors:
            raise errors[0]
        return results
