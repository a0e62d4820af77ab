"""Keep a rank's host threads and page-locked memory on the CPU socket its GPU hangs off.

The drop-in seam (``evolvers.evolve_space`` / ``evolve_time``) and the snapshot path move whole grids between host arrays
and the device every step; with one process per GPU on a two-socket box, a rank whose page-locked arrays sit on the
other socket pays the inter-socket link on every transfer.  ``bind_host_to_gpu`` restricts the calling process to the
CPUs NVML reports as local to the GPU (intersected with what the process may use), before any page-locked memory is
allocated, so that first-touch placement puts the arrays and the staging lanes (``astrea_upload`` / ``astrea_download``,
api.cu ``staged_copy``) on that socket.  It is opt-in: a library does not change its host's affinity on its own;
``bench.py`` calls it for every rank.
"""
import os


def _visible_index(device):
    """Index NVML knows the CUDA device under (CUDA_VISIBLE_DEVICES may renumber or name devices by UUID)."""
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if not vis:
        return device
    entries = [e.strip() for e in vis.split(",") if e.strip()]
    if device >= len(entries):
        return device
    return entries[device]


def bind_host_to_gpu(device=0):
    """Restrict this process to the CPUs local to CUDA device ``device``.  Returns a dict describing what was done
    (``{"bound": False, "why": ...}`` when NVML, the affinity call or a non-trivial answer is unavailable)."""
    if not hasattr(os, "sched_setaffinity"):
        return {"bound": False, "why": "no sched_setaffinity on this platform"}
    try:
        import pynvml
        pynvml.nvmlInit()
        which = _visible_index(device)
        if isinstance(which, str) and not which.isdigit():
            handle = pynvml.nvmlDeviceGetHandleByUUID(which.encode() if hasattr(which, "encode") else which)
        else:
            handle = pynvml.nvmlDeviceGetHandleByIndex(int(which))
        words = (max(os.sched_getaffinity(0) | {os.cpu_count() or 1}) // 64) + 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        local = {64 * w + b for w in range(len(mask)) for b in range(64) if (int(mask[w]) >> b) & 1}
        try:
            numa = int(pynvml.nvmlDeviceGetNumaNodeId(handle))
        except Exception:
            numa = None
    except Exception as exc:       # NVML missing or too old: leave the affinity alone
        return {"bound": False, "why": f"NVML: {exc!r}"}
    allowed = os.sched_getaffinity(0)
    cpus = local & allowed
    if not cpus:
        return {"bound": False, "why": "the GPU's local CPUs are outside this process's cpuset", "numa_node": numa}
    if cpus == allowed:
        return {"bound": False, "why": "every allowed CPU is local to the GPU", "numa_node": numa, "cpus": len(cpus)}
    os.sched_setaffinity(0, cpus)
    return {"bound": True, "numa_node": numa, "cpus": len(cpus), "of": len(allowed)}
