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
