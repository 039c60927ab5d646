"""The caller side of `Physics.update()`: the reference's physics thread, its frame clock
and the snapshot hand-off to the renderer, for a headless driver of the native backend.

  Clock            B/Clock.java        ring of the last n frame times: dt, mean, variance, frame rate
  Loop             B/Loop.java         thread that calls a callback with the (capped) frame time, runs queued
                                       commands between iterations, can be paused / stopped / abandoned
  PhysicsSnapshot  A/PhysicsSnapshot.java   host copy of the particle arrays + settings + type histogram
  Simulation       A/Main.java:245-300,583-604,993-1016  the glue: `update_physics(real_dt)` picks dt and steps, a
                                       `do_once` command takes a snapshot when the consumer asks for the next, and the
                                       "physics not reacting" reset

Everything the reference serialises through `loop.enqueue(...)` (setters, matrix edits, cursor edits,
loading a save) must go through `Loop.enqueue` here too: the native handle is not thread safe, and like the
reference the loop thread is the only one that touches it while the loop runs.
"""
from __future__ import annotations

import math
import threading
import time
from collections import deque
from typing import Callable, Optional

import numpy as np


class Clock:
    """B/Clock.java:5-113.  Times are in milliseconds."""

    def __init__(self, n: int = 20, now: Callable[[], int] = time.perf_counter_ns):
        self._now = now
        self._in: Optional[int] = None
        self._last = [0.0] * n
        self._idx = -1
        self._dt = 0.0
        self._avg = 0.0
        self._var = 0.0

    def tick(self):
        """out() then in(): time between consecutive ticks (:29-34)."""
        if self._in is not None:
            self.out()
        self.enter()

    def enter(self):
        """`in()` in the reference (a Python keyword)."""
        self._in = self._now()

    def out(self):
        """:43-75: record the interval and refresh mean and (n-1) variance over the whole ring, zeros included."""
        if self._in is None:
            raise RuntimeError("Clock.out() was called even though Clock.in() was never called before")
        self._dt = (self._now() - self._in) / 1e6
        n = len(self._last)
        self._idx = (self._idx + 1) % n
        self._last[self._idx] = self._dt
        if n < 2:
            self._avg, self._var = self._last[0], 0.0
        else:
            s = math.fsum(self._last)
            sq = math.fsum(t * t for t in self._last)
            self._avg = s / n
            self._var = (sq - n * self._avg * self._avg) / (n - 1)

    @property
    def avg_dt_millis(self) -> float:
        return self._avg

    @property
    def avg_framerate(self) -> float:
        return 0.0 if self._avg == 0 else 1000.0 / self._avg

    @property
    def standard_deviation(self) -> float:
        return math.sqrt(max(self._var, 0.0))

    @property
    def dt_millis(self) -> float:
        return self._dt

    @property
    def framerate(self) -> float:
        return 0.0 if self._dt == 0 else 1000.0 / self._last[self._idx]


class Loop:
    """B/Loop.java:32-205.

    `kill()` cannot interrupt a Python thread; it clears the run flag and abandons the (daemon) thread, which
    ends after its current callback.  Pair it with `Physics.force_update_stop()` like Main does (A/Main.java:331-335).
    """

    def __init__(self, clock: Optional[Clock] = None):
        self.max_dt = 1.0 / 20.0          # :38; negative = no cap
        self.pause = False                # :43
        self._clock = clock or Clock(60)  # :45
        self._thread: Optional[threading.Thread] = None
        self._run = threading.Event()
        self._queue: deque = deque()
        self._once = None
        self._once_lock = threading.Lock()
        self._lock = threading.RLock()
        self.error: Optional[BaseException] = None

    def enqueue(self, cmd: Callable[[], None]):
        """:72-75: FIFO, drained at the start of the next iteration."""
        self._queue.append(cmd)

    def do_once(self, cmd: Callable[[], None]):
        """:87-89: single slot, the latest command wins."""
        with self._once_lock:
            self._once = cmd

    def start(self, callback: Callable[[float], None]):
        """:98-109"""
        with self._lock:
            if self._thread is not None:
                raise RuntimeError("Loop thread didn't finish properly (wasn't null).")
            self._run.set()
            run = self._run

            def body():
                try:
                    while run.is_set():
                        self.iterate(callback)
                except BaseException as e:  # a failing command ends the loop; keep why
                    self.error = e
                    run.clear()

            self._thread = threading.Thread(target=body, name="plife-loop", daemon=True)
            self._thread.start()

    def iterate(self, callback: Callable[[float], None]):
        """One iteration (:111-122); public so a single-threaded driver or a test can pump the loop by hand."""
        self._clock.tick()
        while True:
            try:
                cmd = self._queue.popleft()
            except IndexError:
                break
            cmd()
        with self._once_lock:
            once, self._once = self._once, None
        if once is not None:
            once()
        if not self.pause:
            callback(self._compute_dt())

    def stop(self, millis: int = 0) -> bool:
        """:139-165: wait for the current iteration; on timeout leave the loop running and return False."""
        with self._lock:
            t = self._thread
            if t is None:
                return True
            if not t.is_alive():
                self._thread = None
                return True
            was = self._run.is_set()
            self._run.clear()
            t.join(None if millis == 0 else millis / 1000.0)
            if t.is_alive():
                if was:
                    self._run.set()
                return False
            self._thread = None
            return True

    def kill(self):
        """:178-189"""
        with self._lock:
            if self._thread is None:
                return
            self._run.clear()
            self._run = threading.Event()   # a later start() must not revive the abandoned thread
            self._thread = None

    @property
    def running(self) -> bool:
        t = self._thread
        return t is not None and t.is_alive()

    def _compute_dt(self) -> float:
        """:191-194.  The reference takes min(maxDt, dt) whatever the sign, so a negative maxDt yields a negative
        dt despite the field's doc comment; kept."""
        return min(self.max_dt, self._clock.dt_millis / 1000.0)

    def get_actual_dt(self) -> float:
        return self._clock.dt_millis / 1000.0

    def get_avg_framerate(self) -> float:
        return self._clock.avg_framerate


class PhysicsSnapshot:
    """A/PhysicsSnapshot.java:8-58.  Arrays are (n, 2) rather than the reference's flat xyz triples (z is always 0).

    `take` reuses pinned-size host arrays when the count is unchanged and fills them straight from the device
    in fp32 (`plife_download_f32`), which is what the renderer uploads anyway (A/Main.java:585-598).
    """

    def __init__(self):
        self.positions = np.zeros((0, 2), np.float32)
        self.velocities = np.zeros((0, 2), np.float32)
        self.types = np.zeros(0, np.int32)
        self.settings = None
        self.particle_count = 0
        self.type_count = np.zeros(0, np.int64)
        self.snapshot_time = 0  # unix ms

    def take(self, physics):
        n = physics.particle_count
        if self.types.shape[0] != n:
            self.positions = np.empty((n, 2), np.float32)
            self.velocities = np.empty((n, 2), np.float32)
            self.types = np.empty(n, np.int32)
        if n:
            physics.native.download_f32(self.positions, self.velocities, self.types)
        self.settings = physics.settings.deep_copy()
        self.particle_count = n
        self.type_count = np.asarray(physics.get_type_count())
        self.snapshot_time = int(time.time() * 1000)


class Simulation:
    """Headless stand-in for the parts of A/Main.java that drive the backend.

    update_physics   A/Main.java:291-295   dt = real frame time when `auto_dt`, else the fixed `dt`
    request_snapshot A/Main.java:583-604   consumer polls `new_snapshot_available`, reads, then asks for the next
    close            A/Main.java:329-337   stop(1000) else kill() + force_update_stop()
    """

    def __init__(self, physics, *, auto_dt: bool = True, dt: float = 0.02):
        self.physics = physics
        self.auto_dt = auto_dt
        self.dt = dt
        self.loop = Loop()
        self.snapshot = PhysicsSnapshot()
        self.new_snapshot_available = threading.Event()
        self.steps = 0
        self.snapshot.take(physics)          # :286-289
        self.new_snapshot_available.set()

    def update_physics(self, real_dt: float):
        self.physics.settings.dt = real_dt if self.auto_dt else self.dt
        self.physics.update()
        self.steps += 1

    def start(self):
        self.loop.start(self.update_physics)

    def request_snapshot(self):
        def take():
            self.snapshot.take(self.physics)
            self.new_snapshot_available.set()
        self.new_snapshot_available.clear()
        self.loop.do_once(take)

    def close(self, millis: int = 1000) -> bool:
        clean = self.loop.stop(millis)
        if not clean:
            self.loop.kill()
            self.physics.force_update_stop()
        return clean

    # -- watchdog (A/Main.java:993-1016) --
    def not_reacting_for(self) -> float:
        """Milliseconds since the last snapshot was taken - the reference's only liveness signal (snapshots are
        requested every frame, so an old one means the physics thread is stuck)."""
        return time.time() * 1000.0 - self.snapshot.snapshot_time

    def reset_if_not_reacting(self, make_physics: Callable[[], object], threshold_ms: float = 3000.0, stop_millis: int = 1000) -> bool:
        """What the "Reset Physics" button does once the physics has not reacted for `threshold_ms` (the app's
        physicsNotReactingThreshold): stop the loop, or abandon it and kill the physics if it does not stop in time,
        then start over with a fresh physics object and a fresh loop.  Returns True if a reset happened."""
        if self.not_reacting_for() <= threshold_ms:
            return False
        if not self.loop.stop(stop_millis):
            # The abandoned loop thread may still be inside the native step on this handle.  The handle is single-owner
            # (include/plife.h): only request_stop may come from another thread, and destroying it under a running call
            # would be a use-after-free.  So: ask it to stop now, free it once the abandoned thread has left.
            old_thread, old_physics = self.loop._thread, self.physics
            self.loop.kill()
            try:
                old_physics.force_update_stop()
            except Exception:  # noqa: BLE001 - a wedged backend must not keep the reset from happening
                pass

            def reap():
                if old_thread is not None:
                    old_thread.join()
                try:
                    old_physics.kill()
                except Exception:  # noqa: BLE001
                    pass
            threading.Thread(target=reap, name="plife-reaper", daemon=True).start()
        self.physics = make_physics()
        self.loop = Loop()
        self.steps = 0
        self.snapshot.take(self.physics)
        self.new_snapshot_available.set()
        self.loop.start(self.update_physics)
        return True
