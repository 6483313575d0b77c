"""besides one process per GPU, "one process, one host thread per device".  All mutable library
state is thread_local (context.hpp): a host thread is a rank — it binds to one GPU, owns the handles it creates and
its own constant-table tag, buffer pools, model cache and NCCL communicator.  One thread per device: the
constant-memory tables are per device, so a second thread binding to an owned device is refused loudly."""
import threading

import numpy as np
import pytest

from quickchem_b200 import synth

pytestmark = [pytest.mark.gpu]


def test_second_thread_per_device_rules(capi, oracle, small_model_path):
    x = synth.quick_features(synth.raw_fields(6))
    ref = oracle.Model(small_model_path).predict(x)
    b_main = capi.Booster(small_model_path)  # binds the main thread to its device (0 unless LOCAL_RANK says otherwise)
    assert np.array_equal(b_main.predict(capi.DMatrix(x)).view(np.uint32), ref.view(np.uint32))
    ndev = capi.device_count()
    out = {}

    def worker(device):
        try:
            L = capi.lib()
            # a handle of another thread is not a handle here
            if L.qcoh_booster_get_info(b_main.handle, capi.C.byref(capi.BoosterInfo())) == 0:
                out["foreign"] = "a foreign handle was accepted"
                return
            rc = L.qcoh_set_device(device)
            if rc != 0:
                out["bind_error"] = capi.last_error()
                return
            b = capi.Booster(small_model_path)
            out["pred"] = b.predict(capi.DMatrix(x))
            out["kernel"] = capi.last_predict_kernel()
            out["launches"] = capi.launch_count()
            b.free()
        except Exception as e:  # noqa: BLE001
            out["exc"] = repr(e)

    # the main thread's device is taken
    t = threading.Thread(target=worker, args=(0,))
    t.start()
    t.join()
    assert "exc" not in out and "foreign" not in out, out
    assert "already driven by another host thread" in out.get("bind_error", ""), out
    if ndev >= 2:  # a second GPU: the second thread is a second rank of the same process
        out.clear()
        n_main = capi.launch_count()
        t = threading.Thread(target=worker, args=(1,))
        t.start()
        t.join()
        assert "exc" not in out and "bind_error" not in out, out
        assert np.array_equal(out["pred"].view(np.uint32), ref.view(np.uint32))
        assert out["launches"] > 0 and capi.launch_count() == n_main  # counters are per thread
    # the main thread is unaffected
    assert np.array_equal(b_main.predict(capi.DMatrix(x)).view(np.uint32), ref.view(np.uint32))
