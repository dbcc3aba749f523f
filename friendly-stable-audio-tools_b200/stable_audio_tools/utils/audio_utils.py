"""Audio output helpers.

``float_to_int16_audio`` has the semantics of the reference's ``utils/audio_utils.py:22-27``
(peak-normalise only when the peak exceeds 1, or always with ``maximize``; 16-bit PCM on the host).
``stream_decode_int16`` is the pipelined form the generation scripts want for long outputs
(SURVEY 8f-3): latents are decoded one sample at a time and the int16 conversion + device-to-host copy
of sample i run on a side stream while the decoder already works on sample i + 1.
"""
import torch


def float_to_int16_audio(x: torch.Tensor, maximize: bool = False):
    peak = float(x.abs().max())
    divisor = peak if maximize else max(peak, 1.0)
    return (x / divisor * 32767).to(torch.int16).cpu()      # same operation order as the reference


@torch.no_grad()
def stream_decode_int16(decode_fn, latents: torch.Tensor, maximize: bool = False):
    """Yield one int16 CPU tensor [channels, samples] per latent sample, in order.

    ``decode_fn(z[1, C, L]) -> audio[1, channels, T]`` is e.g. ``model.pretransform.decode``.  Every sample is
    normalised on its own (like calling ``float_to_int16_audio`` per file, ``generate.py:142-151`` of the
    reference).  Three pinned host buffers are reused in rotation: sample k is handed out while sample k + 1 is in
    flight; the next advance (which hands out k + 1) enqueues sample k + 2 into another buffer, and only the advance
    after that reuses sample k's buffer.  So a yielded tensor stays valid across ONE further advance of the
    generator (a one-deep writer queue is safe) and must be consumed or copied before the second.
    """
    if not latents.is_cuda:
        raise RuntimeError("stream_decode_int16 needs CUDA latents (the decoder has no CPU path)")
    main = torch.cuda.current_stream(latents.device)
    side = torch.cuda.Stream(device=latents.device)
    n_slots = 3
    pinned, done, pending = [None] * n_slots, [None] * n_slots, None
    for i in range(latents.shape[0]):
        audio = decode_fn(latents[i:i + 1])[0]
        ready = torch.cuda.Event()
        ready.record(main)
        slot = i % n_slots
        if done[slot] is not None:
            done[slot].synchronize()                    # copy of sample i - 3; the consumer has had one full advance since it was handed out
        with torch.cuda.stream(side):
            side.wait_event(ready)
            peak = audio.abs().max()
            div = peak if maximize else torch.clamp(peak, min=1.0)
            # torch divides a CUDA tensor by a host scalar as a multiplication with its fp32 reciprocal; do the
            # same here so the streamed samples equal float_to_int16_audio's
            pcm = (audio * (1.0 / div) * 32767).to(torch.int16)
            if pinned[slot] is None or pinned[slot].shape != pcm.shape:
                pinned[slot] = torch.empty(pcm.shape, dtype=torch.int16, pin_memory=True)
            pinned[slot].copy_(pcm, non_blocking=True)
            audio.record_stream(side)
            done[slot] = torch.cuda.Event()
            done[slot].record(side)
        if pending is not None:                         # hand out the previous sample while this one is in flight
            done[pending].synchronize()
            yield pinned[pending]
        pending = slot
    if pending is not None:
        done[pending].synchronize()
        yield pinned[pending]
