"""Construction timings on the GPU box: wt_huff / wt_int bit planes on the device (wt_build.cu)
against this library's multi-threaded host fill (SDSLGPU_HOST_WT=1) and the reference's single-thread constructor,
the FM-index end to end, and egress (sdslgpu_serialize -> the reference's store_to_file bytes).  One JSON line each."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
orc = ge.load_oracle()


def timed(fn):
    t0 = time.perf_counter()
    r = fn()
    return r, time.perf_counter() - t0


def main():
    logn = int(sys.argv[1]) if len(sys.argv) > 1 else 28
    with_ref = (len(sys.argv) > 2 and sys.argv[2] == "ref") and orc.ref_available()
    rng = np.random.default_rng(42)
    text = rng.integers(1, 256, 1 << logn, dtype=np.uint8).tobytes()
    pkg.WtHuff(text[: 1 << 16]).close()  # context + module load outside the timings
    os.environ.pop("SDSLGPU_HOST_WT", None)
    wt, dev_s = timed(lambda: pkg.WtHuff(text))
    blob, ser_s = timed(wt.serialize)
    wt.close()
    os.environ["SDSLGPU_HOST_WT"] = "1"
    wt2, host_s = timed(lambda: pkg.WtHuff(text))
    same = wt2.serialize() == blob
    wt2.close()
    os.environ.pop("SDSLGPU_HOST_WT", None)
    line = {"op": "wt_huff construct", "symbols": 1 << logn, "device_planes_s": dev_s, "host_planes_s": host_s, "identical_blobs": same,
            "serialize_s": ser_s, "blob_bytes": len(blob)}
    if with_ref:
        r, ref_s = timed(lambda: orc.Ref().wt_huff(text))
        line["reference_s"] = ref_s
        line["blob_equals_reference"] = r.serialize() == blob
    print(json.dumps(line), flush=True)

    seq = rng.integers(0, 1 << 20, 1 << (logn - 2), dtype=np.uint64)
    wi, dev_s = timed(lambda: pkg.WtInt(seq))
    bi = wi.serialize()
    wi.close()
    os.environ["SDSLGPU_HOST_WT"] = "1"
    wi2, host_s = timed(lambda: pkg.WtInt(seq))
    same = wi2.serialize() == bi
    wi2.close()
    os.environ.pop("SDSLGPU_HOST_WT", None)
    print(json.dumps({"op": "wt_int construct", "symbols": len(seq), "levels": 20, "device_planes_s": dev_s, "host_planes_s": host_s,
                      "identical_blobs": same}), flush=True)

    csa, dev_s = timed(lambda: pkg.CsaWt(text))
    cb, ser_s = timed(csa.serialize)
    csa.close()
    print(json.dumps({"op": "csa_wt construct (suffix array, BWT, samples, wt_huff, occurrence bitmaps)", "symbols": 1 << logn, "device_s": dev_s,
                      "serialize_s": ser_s, "blob_bytes": len(cb)}), flush=True)


if __name__ == "__main__":
    main()
