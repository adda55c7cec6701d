"""Host side of the device-driven frame pipeline shared by both colour back-ends (gauss_render.GaussPythonRenderer,
g2pc.rasterizer.GaussianRasterizer).

A "frame" is one camera.  All kernels of a frame are enqueued without waiting; the sizes later kernels need live in a
device-side header (include/g2pc.h, G2PC_HDR_*).  A frame that does not fit the host's buffers sets the sticky POISON word
on the device: every later kernel of that and the following frames is a no-op.  The host copies the 64-byte header of
every frame to pinned memory asynchronously and looks at it when the copy has landed (or when a getter calls flush()):
on poison it grows what was too small, clears the word and replays the skipped frames in order — so the per-Gaussian
accumulators see the cameras in exactly the reference's order (gauss_to_pc.py:437-454) and no call ever waits for a
count.  The reference synchronises the device several times per camera (rasterizer_impl.cu:289 blocking D2H,
auxiliary.h:178-185 CHECK_CUDA after every stage).
"""
import torch

from . import capi


class FrameQueue:
    """Mixin.  The owner provides: self.device, self._hdr (device int32 header), self._enqueue(camera, frame, index),
    self._fix(header_list) (grow buffers / tables for the failure recorded in the header) and self._confirm(header_list)."""

    def _init_frames(self):
        self._frame = 0
        self._pending = []   # (frame, camera, camera_index, pinned header, event)
        self._hdr_pool = []
        self.replays = 0
        self.async_mode = False

    def _record(self, frame, camera, camera_index):
        hdr = self._hdr_pool.pop() if self._hdr_pool else torch.zeros((capi.HDR_WORDS,), dtype=torch.int32).pin_memory()
        hdr.copy_(self._hdr, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._pending.append((frame, camera, camera_index, hdr, ev))

    def _submit(self, camera, camera_index=None):
        frame = self._frame
        self._frame += 1
        camera_index = frame if camera_index is None else camera_index
        out = self._enqueue(camera, frame, camera_index)
        self._record(frame, camera, camera_index)
        if not self.async_mode:
            self.flush()
        else:
            self._poll(block_if_more_than=8)
        return out

    def _poll(self, block_if_more_than=None):
        while self._pending:
            frame, camera, cidx, hdr, ev = self._pending[0]
            if not ev.query():
                if block_if_more_than is None or len(self._pending) <= block_if_more_than:
                    return
                ev.synchronize()
            h = hdr.tolist()
            if h[capi.HDR_POISON]:
                self._recover(h)
                continue
            self._confirm(h)
            self._hdr_pool.append(hdr)
            self._pending.pop(0)

    def flush(self):
        """Wait for every enqueued frame and replay the ones a poisoned header skipped."""
        while self._pending:
            self._pending[-1][4].synchronize()
            self._poll(block_if_more_than=0)

    def _recover(self, h):
        torch.cuda.current_stream(self.device).synchronize()
        failed = h[capi.HDR_POISON] - 1
        todo = [p for p in self._pending if p[0] >= failed]
        self._pending = [p for p in self._pending if p[0] < failed]
        self._fix(h)
        self._hdr.zero_()
        self.replays += 1
        for (frame, camera, cidx, hdr, ev) in todo:
            self._hdr_pool.append(hdr)
            self._enqueue(camera, frame, cidx)
            self._record(frame, camera, cidx)
