"""A/B timing of library builds on one box: ab_time.py <log_n,...> <lib or '-'> [<lib> ...] (development tool)."""
import ctypes as C, json, os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, os.path.join(HERE, ".."))
    import fftb200_loader
    F = fftb200_loader.load(); L = F.lib
    F.require_gpu()
    for lg in [int(a) for a in sys.argv[2].split(",")]:
        n = 1 << lg; batch = (1 << 28) >> lg; tot = n * batch
        m_in = L.fft_gpu_alloc(tot); m_out = L.fft_gpu_alloc(tot)
        assert m_in and m_out, L.fftb200_last_error()
        L.fftb200_fill_splitmix(L.fftb200_devptr_of(m_in), 43, 0, tot)
        plan = L.fft_gpu_plan_1d(n, batch, -1); eng = L.fftb200_engine_of(plan)
        din, dout = L.fftb200_devptr_of(m_in), L.fftb200_devptr_of(m_out)
        ts = []; ms = C.c_float()
        for i in range(int(sys.argv[3])):
            rc = L.fftb200_timer_start(eng) or L.fftb200_plan_exec_async(eng, din, dout) or L.fftb200_timer_stop(eng, C.byref(ms))
            if rc:
                print("FAILED at rep", i, L.fftb200_last_error()); sys.exit(1)
            if i >= 3: ts.append(ms.value)
        ts.sort()
        print(json.dumps({"lib": os.path.basename(os.environ.get("FFTB200_LIB", "default")), "log_n": lg, "best": round(ts[0], 4), "med": round(ts[len(ts) // 2], 4),
                          "p90": round(ts[len(ts) * 9 // 10], 4), "TBps_best": round(32 * tot / ts[0] * 1e-9, 3)}), flush=True)
        L.fft_gpu_destroy_plan(plan); L.fft_gpu_free(m_in); L.fft_gpu_free(m_out)
    sys.exit(0)
logs, libs = sys.argv[1], sys.argv[2:]
reps = os.environ.get("AB_REPS", "40")
for rnd in range(2):
    for lib in libs:
        env = dict(os.environ)
        if lib != "-": env["FFTB200_LIB"] = os.path.join(HERE, "..", "fft-implementation-in-c_b200", "lib", lib)
        subprocess.run([sys.executable, __file__, "--child", logs, reps], env=env)
