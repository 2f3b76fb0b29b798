"""Layer-split pipeline schedule (host logic of `bench.py --gpus N`, mirrors LLAMA_SPLIT_MODE_LAYER:
/root/reference/llama.cpp/src/llama-model.cpp:1917-1958 layer→device, ggml-backend.cpp:1360-1393 hand-off).

Rank r owns a contiguous range of layers.  `world` sequences are in flight, one per pipeline slot: at tick t rank r
works on sequence (t - r) mod world.  Per tick a rank receives one tensor (rank 0: the token sampled by the last
rank `world` ticks ago; others: the hidden state of rank r-1), computes, and sends one tensor (last rank: the
sampled token to rank 0; others: the hidden state to r+1).  Pure point-to-point; there is no collective on this path.
"""


def layer_range(rank, world, n_layer):
    return rank * n_layer // world, (rank + 1) * n_layer // world


def tick_plan(rank, world, t, total_ticks):
    """what rank `rank` does at tick t: None while the pipeline fills, else dict(seq, recv_from, send_to)"""
    if t < rank:
        return None
    seq = (t - rank) % world
    first, last = rank == 0, rank == world - 1
    if first:
        recv_from = world - 1 if (t >= world and world > 1) else None      # token of this sequence, sampled `world` ticks ago
    else:
        recv_from = rank - 1
    if last:
        # the token sampled at tick t is consumed by rank 0 at tick t + 1 (world == 1: kept locally)
        send_to = 0 if (world > 1 and t + 1 < total_ticks and t + 1 >= world) else None
    else:
        send_to = rank + 1 if t + 1 < total_ticks else None              # rank r+1 consumes it at tick t + 1
    return dict(seq=seq, recv_from=recv_from, send_to=send_to)


def run(rank, world, total_ticks, stage, recv, send):
    """drive one rank: stage(seq, inp, t) -> out; recv(src) -> tensor (blocking); send(tensor, dst) -> handle with
    .wait() (NON-blocking: rank 0 primes `world` sequences before it ever receives, so a rendezvous send would
    deadlock against the last rank's token send).  Returns the outputs of the last rank."""
    outs, pending = [], None
    for t in range(total_ticks):
        p = tick_plan(rank, world, t, total_ticks)
        if p is None:
            continue
        inp = recv(p["recv_from"]) if p["recv_from"] is not None else None
        out = stage(p["seq"], inp, t)
        if pending is not None:
            pending.wait()
            pending = None
        if p["send_to"] is not None:
            pending = send(out, p["send_to"])
        if rank == world - 1:
            outs.append((t, p["seq"], out))
    if pending is not None:
        pending.wait()
    return outs
