"""FP64 pipe microbenchmarks: instruction-issue rate (as 'TFLOP/s if every instruction were a DFMA') per form.
The per-operand-count table the cost model uses comes from tools/proto/regbank.cu (stand-alone, controls the register numbers);
this script is the quick check through the library (rbq_fp64_peak with RBQ_PEAK_MODE)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rbqoc_b200 as rbq

eng = rbq.Engine(0)
names = {0: "DFMA reg,reg,reg (roofline denominator)", 1: "DMUL", 2: "DADD", 3: "DFMA with immediate addend",
         4: "1 DMUL : 3 DFMA", 5: "DFMA three distinct moving regs", 6: "DFMA negated + immediate", 7: "DFMA one reusable + two fresh regs",
         8: "DFMA three fresh regs, accumulate in place"}
for mode, name in names.items():
    eng.set_option("RBQ_PEAK_MODE", mode)
    v = eng.fp64_peak(4000)
    print(f"mode {mode}: {v:7.2f} 'TFLOP/s'  {v / 37.2 * 100:5.1f}% of DFMA issue peak   {name}", flush=True)

eng.set_option("RBQ_PEAK_MODE", 9)
v = eng.fp64_peak(20000)          # "TFLOP/s" = sm_count*32 threads * 2 flop / (latency in seconds)
lat_s = eng._lib and (148 * 32 * 2) / (v * 1e12)
print(f"dependent DFMA chain, 1 warp/SM: {lat_s * 1.965e9:.2f} cycles per DFMA at 1965 MHz (issue-to-issue latency)")
