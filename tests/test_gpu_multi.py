"""Real multi-GPU parity (NCCL, one process per GPU): skipped on a box with fewer than two GPUs. The multi-process
host logic is additionally covered on CPU over gloo (tests/test_sharded_cpu.py) and by single-GPU emulations
(tests/test_gpu_parity.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_and_replicated_search_equal_the_oracle_on_real_gpus(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, have {torch.cuda.device_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0 and f"MULTI_GPU_PARITY_OK {world}" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.parametrize("n_dev", [2, 4])
def test_one_handle_over_several_devices(n_dev):
    """kmer_b200_config.device_ids / n_devices: one process, one handle, the index replicated over the devices from
    key-range parts (peer copies), a host batch striped over them. Must equal the oracle and the one-device index."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < n_dev:
        pytest.skip(f"needs {n_dev} GPUs, have {torch.cuda.device_count()}")
    import kmer_index_b200 as kb
    from conftest import assert_results_equal
    from kmer_index_b200 import synth
    from oracle import bindings
    bindings.build()
    # heavy = a 70 000-element bucket: such an index is replicated whole on every device; without one the positions stay
    # in the devices' parts (read over NVLink) and only the directory is made whole everywhere
    for sigma, ks, n, m_lo, m_hi, heavy in [(4, [16], 1_500_000, 16, 64, True), (4, [16], 1_500_000, 16, 64, False),
                                            (4, [12], 3_000_000, 5, 60, False), (4, [5, 7, 9, 11, 13], 300_000, 4, 40, True),
                                            (27, [8], 200_000, 3, 30, True)]:
        text = synth.random_text(n, sigma, 41)
        if heavy:
            text[70_000:140_000] = 0
        q, off = synth.stress_queries(text, 30_000, m_lo, m_hi, sigma, 42)
        with bindings.Oracle(text, sigma, ks) as o:
            want = o.search(q, off)
            want_csr = [o.element(e) for e in range(len(ks))]
        with kb.KmerIndex(text, sigma, ks, devices=list(range(n_dev))) as ix:
            for attempt in range(2):
                assert_results_equal(ix.search_batch(q, off).as_tuple(), want, label=f"{n_dev} devices {ks}/{attempt}")
            for e in range(len(ks)):
                if heavy:
                    h, p = ix.element_arrays(e)
                    assert np.array_equal(p, want_csr[e][1]) and np.array_equal(h.astype(np.uint64), want_csr[e][0])
                else:
                    with pytest.raises(kb.KmerB200Error):      # no device holds the whole position array
                        ix.element_arrays(e)
            assert ix.device_bytes > 0
            with pytest.raises(kb.KmerB200Error):          # device-pointer entry points need a single-device handle
                ix.count_batch_device(0, 0, 0, 1)
