export GSM_LIB_PATH=$PWD/build/libgsm_dev.so
python tools/tune.py C4 125000 2000 128x3 2>&1 | tail -1 | cut -c60-140
python - <<'PY'
import os,sys,json
sys.path.insert(0,'tests')
from gammaspikemodel_b200 import _lib
_lib.LIB_PATH=os.environ["GSM_LIB_PATH"]
import torch
from gammaspikemodel_b200 import GsmEngine, synthetic, abi
from helpers import host_arrays, run_capi, run_oracle, assert_results_match
for kind,T,n in (("C4",64,2000),("C3",64,1000),("C2",64,300)):
    b=synthetic.make_batch(kind,T,n); a=host_arrays(b)
    assert_results_match(run_capi(a,b.flags,abi.STUBS_NOSTUB), run_oracle(a,b.flags,abi.STUBS_NOSTUB), what=kind)
print("mixture parity ok")
dev=torch.device("cuda",0)
for kind,trees,taxa in (("C3",20000,1000),("C4",20000,1000)):
    b = synthetic.make_batch_chunked(kind, trees, taxa, device=dev, chunk=16384)
    out = torch.empty((b.n_trees, abi.NOUT), dtype=torch.float64, device=dev)
    st = torch.cuda.Stream(); torch.cuda.set_stream(st)
    eng = GsmEngine(device=0); eng.set_mode(abi.STUBS_NOSTUB, b.flags); eng.bind_batch(b); eng.set_option("block",128); eng.set_option("minb",3)
    for _ in range(2): eng.eval_device(out, None, None, st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): eng.eval_device(out, None, None, st.cuda_stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(kind, "mixture", ms, b.n_trees*b.n_nodes/ms*1e3, float(torch.isfinite(out[:,3]).double().mean()))
PY
