"""astrea_b200.hostbind: the rank's CPU affinity follows what NVML reports as local to its GPU (NVML faked here)."""
import os
import sys
import types

import pytest

from astrea_b200 import hostbind


def _fake_nvml(mask_words, numa=1, seen=None):
    mod = types.ModuleType("pynvml")
    mod.nvmlInit = lambda: None

    def by_index(i):
        if seen is not None:
            seen.append(("index", i))
        return ("h", i)

    def by_uuid(u):
        if seen is not None:
            seen.append(("uuid", u))
        return ("h", u)
    mod.nvmlDeviceGetHandleByIndex = by_index
    mod.nvmlDeviceGetHandleByUUID = by_uuid
    mod.nvmlDeviceGetCpuAffinity = lambda h, n: list(mask_words)[:n] + [0] * max(0, n - len(mask_words))
    mod.nvmlDeviceGetNumaNodeId = lambda h: numa
    return mod


@pytest.mark.skipif(not hasattr(os, "sched_setaffinity"), reason="no sched_setaffinity")
def test_binds_to_the_gpus_local_cpus(monkeypatch):
    calls, seen = [], []
    monkeypatch.setitem(sys.modules, "pynvml", _fake_nvml([0b1111], seen=seen))      # CPUs 0-3 are local to the GPU
    monkeypatch.setattr(os, "sched_getaffinity", lambda pid: {0, 1, 2, 3, 4, 5, 6, 7})
    monkeypatch.setattr(os, "sched_setaffinity", lambda pid, cpus: calls.append(set(cpus)))
    monkeypatch.setenv("CUDA_VISIBLE_DEVICES", "5,6")
    info = hostbind.bind_host_to_gpu(1)
    assert info == {"bound": True, "numa_node": 1, "cpus": 4, "of": 8} and calls == [{0, 1, 2, 3}]
    assert seen == [("index", 6)]                       # CUDA device 1 is NVML device 6 under this CUDA_VISIBLE_DEVICES
    # every allowed CPU local (one socket, or a cpuset inside it): nothing to do
    calls.clear()
    monkeypatch.setattr(os, "sched_getaffinity", lambda pid: {1, 2})
    info = hostbind.bind_host_to_gpu(0)
    assert info["bound"] is False and "local" in info["why"] and calls == []
    # the GPU's CPUs are outside the cpuset: leave the affinity alone
    monkeypatch.setattr(os, "sched_getaffinity", lambda pid: {8, 9})
    info = hostbind.bind_host_to_gpu(0)
    assert info["bound"] is False and "outside" in info["why"] and calls == []


def test_without_nvml_nothing_changes(monkeypatch):
    calls = []
    broken = types.ModuleType("pynvml")

    def boom():
        raise RuntimeError("no driver")
    broken.nvmlInit = boom
    monkeypatch.setitem(sys.modules, "pynvml", broken)
    if hasattr(os, "sched_setaffinity"):
        monkeypatch.setattr(os, "sched_setaffinity", lambda pid, cpus: calls.append(cpus))
    info = hostbind.bind_host_to_gpu(0)
    assert info["bound"] is False and calls == []
