mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_keygen.py -m gpu -x -q -s > gpurun_out/c13_pytest_kg.txt 2>&1; tail -30 gpurun_out/c13_pytest_kg.txt
python - <<'PY' > gpurun_out/c13_keygen_time.txt 2>&1
import importlib, time, sys, os
sys.path.insert(0, os.getcwd())
T = importlib.import_module("go-tfhe_b200")
for name in ("128", "uint5"):
    P = T.params.get(name); sk = T.key.NewSecretKey(P, 1)
    t = time.time(); ck = T.cloudkey.NewCloudKey(sk, 2); ctx = ck.engine(0); t_host = time.time() - t; ck.close()
    t = time.time(); ck = T.cloudkey.NewCloudKeyOnDevice(sk, 2, export=False); t_dev = time.time() - t
    t = time.time(); ck2 = T.cloudkey.NewCloudKeyOnDevice(sk, 3, export=True); t_dev_x = time.time() - t
    print(name, "host keygen + upload s %.2f" % t_host, " device keygen s %.3f" % t_dev, " device keygen + export s %.3f" % t_dev_x, flush=True)
    ck.close(); ck2.close()
PY
cat gpurun_out/c13_keygen_time.txt
