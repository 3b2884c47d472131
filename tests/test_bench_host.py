"""Host-side logic of bench.py that runs without a GPU: the CPU-list parser and the GPU -> CPU affinity lookup
(sysfs numa_node first, `nvidia-smi topo -m` as the fallback on virtualised boxes)."""
import importlib.util
import os
import sys
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def bench():
    spec = importlib.util.spec_from_file_location('bench_under_test', os.path.join(ROOT, 'bench.py'))
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ['bench.py']
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


def test_cpulist(bench):
    assert bench._cpulist('0-3,8-9, 12') == {0, 1, 2, 3, 8, 9, 12}
    assert bench._cpulist('5') == {5}
    assert bench._cpulist('') == set()
    assert bench._cpulist('N/A') == set()


TOPO = ('\t\x1b[4mGPU0\tGPU1\tNIC0\tCPU Affinity\tNUMA Affinity\tGPU NUMA ID\x1b[0m\n'
        'GPU0\t X \tNV18\tSYS\t0-1\t0\t\tN/A\n'
        'GPU1\tNV18\t X \tSYS\t2-3\t1\t\tN/A\n'
        'NIC0\tSYS\tSYS\t X \t\t\t\t\n')


def _fake_run(topo):
    def run(cmd, **kw):
        if 'topo' in cmd:
            return types.SimpleNamespace(stdout=topo, returncode=0)
        return types.SimpleNamespace(stdout='00000000:1B:00.0\n', returncode=0)     # pci.bus_id query
    return run


def test_numa_binding_falls_back_to_topo(bench, monkeypatch):
    """No sysfs entry for the GPU (as in a container): the CPU-affinity column of `nvidia-smi topo -m` is used, restricted
    to the CPUs this process may run on; the note says what happened."""
    bound = {}
    monkeypatch.setattr(bench.subprocess, 'run', _fake_run(TOPO))
    monkeypatch.setattr(bench.os, 'sched_getaffinity', lambda pid: {0, 1, 2, 3})
    monkeypatch.setattr(bench.os, 'sched_setaffinity', lambda pid, cpus: bound.setdefault('cpus', set(cpus)))
    node = bench.bind_to_gpu_numa(1)
    assert bound['cpus'] == {2, 3}
    assert node == 1
    assert 'nvidia-smi topo CPU affinity 2-3' in bench.NUMA_NOTE


def test_numa_binding_leaves_one_affinity_set_alone(bench, monkeypatch):
    """Virtualised boxes report one CPU set for every GPU: nothing to narrow, nothing bound."""
    one_set = TOPO.replace('0-1', '0-3').replace('2-3', '0-3')
    called = []
    monkeypatch.setattr(bench.subprocess, 'run', _fake_run(one_set))
    monkeypatch.setattr(bench.os, 'sched_getaffinity', lambda pid: {0, 1, 2, 3})
    monkeypatch.setattr(bench.os, 'sched_setaffinity', lambda pid, cpus: called.append(cpus))
    assert bench.bind_to_gpu_numa(0) is None
    assert called == []
    assert 'does not narrow' in bench.NUMA_NOTE
