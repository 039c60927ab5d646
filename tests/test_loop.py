"""Clock / Loop / snapshot hand-off (B/Clock.java, B/Loop.java, A/Main.java's use of them) — host logic, no GPU."""
import threading
import time

import numpy as np
import pytest

from plife.loop import Clock, Loop


class FakeTime:
    def __init__(self):
        self.ns = 1_000_000_000

    def __call__(self):
        return self.ns

    def advance_ms(self, ms):
        self.ns += int(ms * 1e6)


def test_clock_statistics_follow_the_reference():
    ft = FakeTime()
    c = Clock(4, now=ft)
    with pytest.raises(RuntimeError):
        c.out()
    c.tick()                                    # first tick only arms the clock
    assert c.dt_millis == 0 and c.avg_framerate == 0 and c.framerate == 0
    for ms in (10, 20, 30):
        ft.advance_ms(ms)
        c.tick()
    # the ring still holds one zero: the reference averages over all n slots from the start
    assert c.dt_millis == 30 and c.avg_dt_millis == pytest.approx(15.0)
    assert c.standard_deviation == pytest.approx(np.std([10, 20, 30, 0], ddof=1))
    assert c.framerate == pytest.approx(1000 / 30) and c.avg_framerate == pytest.approx(1000 / 15)
    ft.advance_ms(40)
    c.tick()
    ft.advance_ms(50)
    c.tick()                                    # overwrites the oldest (10)
    assert c.avg_dt_millis == pytest.approx(np.mean([50, 20, 30, 40]))
    one = Clock(1, now=ft)
    one.enter(); ft.advance_ms(5); one.out()
    assert one.avg_dt_millis == 5 and one.standard_deviation == 0


def test_loop_iteration_order_and_dt_cap():
    ft = FakeTime()
    loop = Loop(Clock(60, now=ft))
    log = []
    loop.enqueue(lambda: log.append("a"))
    loop.enqueue(lambda: log.append("b"))
    loop.do_once(lambda: log.append("x"))
    loop.do_once(lambda: log.append("once"))    # replaces the previous one
    loop.iterate(lambda dt: log.append(("cb", dt)))
    assert log == ["a", "b", "once", ("cb", 0.0)]
    ft.advance_ms(16)
    loop.iterate(lambda dt: log.append(("cb", dt)))
    assert log[-1] == ("cb", pytest.approx(0.016)) and len(log) == 5   # queue and once slot are empty now
    ft.advance_ms(500)
    loop.iterate(lambda dt: log.append(("cb", dt)))
    assert log[-1] == ("cb", 1 / 20) and loop.get_actual_dt() == pytest.approx(0.5)   # capped for the callback only
    loop.pause = True
    loop.enqueue(lambda: log.append("while paused"))
    ft.advance_ms(1)
    loop.iterate(lambda dt: log.append(("cb", dt)))
    assert log[-1] == "while paused"            # commands still run, the callback does not


def test_loop_thread_start_stop_and_commands_run_on_the_loop_thread():
    loop = Loop()
    seen = []
    n = [0]

    def cb(dt):
        n[0] += 1
        time.sleep(0.001)

    loop.start(cb)
    with pytest.raises(RuntimeError):
        loop.start(cb)
    done = threading.Event()
    loop.enqueue(lambda: (seen.append(threading.current_thread().name), done.set()))
    assert done.wait(2.0) and seen == ["plife-loop"]
    assert loop.stop(2000) and not loop.running and n[0] > 0
    assert loop.stop() is True                  # already stopped
    count = n[0]
    time.sleep(0.01)
    assert n[0] == count
    loop.start(cb)                              # restartable after a clean stop
    assert loop.stop(2000)


def test_loop_stop_timeout_leaves_it_running_then_kill_abandons_it():
    loop = Loop()
    release = threading.Event()
    entered = threading.Event()
    calls = [0]

    def cb(dt):
        calls[0] += 1
        entered.set()
        release.wait(5.0)

    loop.start(cb)
    assert entered.wait(2.0)
    assert loop.stop(50) is False and loop.running          # iteration did not finish in time: nothing changed
    loop.kill()
    assert not loop.running
    release.set()
    time.sleep(0.05)
    assert calls[0] == 1                                    # the abandoned thread exits after its callback
    loop.start(lambda dt: time.sleep(0.001))                # and a new loop can start
    assert loop.stop(2000)


def test_failing_command_stops_the_loop_and_keeps_the_error():
    loop = Loop()
    loop.enqueue(lambda: 1 / 0)
    loop.start(lambda dt: time.sleep(0.001))
    for _ in range(200):
        if loop.error is not None:
            break
        time.sleep(0.005)
    assert isinstance(loop.error, ZeroDivisionError)
    assert loop.stop(1000)


class _FakeNative:
    def download_f32(self, *a):
        pass


class _FakePhysics:
    """Stands in for plife.Physics where no GPU exists: counts updates, can be made to hang."""

    def __init__(self, hang=None):
        from plife import PhysicsSettings
        self.settings = PhysicsSettings()
        self.native = _FakeNative()
        self.particle_count = 0
        self.updates = 0
        self.killed = False
        self.hang = hang

    def update(self):
        self.updates += 1
        if self.hang is not None:
            self.hang.wait(10.0)
        time.sleep(0.001)

    def get_type_count(self):
        return np.zeros(6, np.int64)

    def force_update_stop(self):
        pass

    def kill(self):
        self.killed = True


def test_simulation_dt_policy_snapshots_and_watchdog():
    """A/Main.java:291-295 (dt policy), :583-604 (snapshot hand-off), :993-1016 (reset of a physics that stopped reacting)."""
    from plife.loop import Simulation
    p = _FakePhysics()
    sim = Simulation(p, auto_dt=False, dt=0.0125)
    sim.update_physics(0.5)
    assert p.settings.dt == 0.0125
    sim.auto_dt = True
    sim.update_physics(0.031)
    assert p.settings.dt == 0.031 and sim.steps == 2
    sim.start()
    sim.request_snapshot()
    assert sim.new_snapshot_available.wait(5.0)
    assert sim.reset_if_not_reacting(lambda: _FakePhysics(), threshold_ms=3000) is False     # fresh snapshot: alive
    assert sim.close(2000)
    # a physics that hangs inside update(): no new snapshots, stop() times out, the loop is abandoned and everything restarts
    gate = threading.Event()
    stuck = _FakePhysics(hang=gate)
    sim = Simulation(stuck, auto_dt=False)
    sim.start()
    for _ in range(200):
        if stuck.updates:
            break
        time.sleep(0.005)
    sim.snapshot.snapshot_time -= 10_000                      # as if the last snapshot were ten seconds old
    fresh = []
    assert sim.reset_if_not_reacting(lambda: fresh.append(_FakePhysics()) or fresh[-1], threshold_ms=3000, stop_millis=50) is True
    # the stuck physics is NOT freed while its abandoned thread may still be inside the native step (single-owner handle):
    # it is asked to stop at once and killed only after that thread has left
    assert not stuck.killed and sim.physics is fresh[0] and sim.not_reacting_for() < 3000
    for _ in range(200):
        if fresh[0].updates > 3:
            break
        time.sleep(0.005)
    assert fresh[0].updates > 3                                # the new loop runs the new physics
    gate.set()                                                  # let the abandoned thread finish
    for _ in range(400):
        if stuck.killed:
            break
        time.sleep(0.005)
    assert stuck.killed                                         # ... and now the old handle is released
    assert sim.close(2000)
