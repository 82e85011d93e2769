| quantity | round 1 | round 2 |
|---|---|---|
| `value` (device-resident inputs, 64 Mi particles, 1 GPU) | 1.17e+08 /s, 573 ms | **@V1@ particle-updates/s, @MS1@ ms/step** |
| `e2e` (host buffers in/out every step) | 9.85e+07 /s | **@E1@ /s** (4.8 GB H2D + 9.9 GB D2H per step; ~110 ms of PCIe time is not overlappable with this API: the upload before the first sort, and a, atree, gpot after the last gravity kernel) |
| device-resident main loop, one global step | 1.20e+08 /s, 557 ms | @R1@ /s, @RMS1@ ms/step |
| reference arm (unmodified GANDALF, OpenMP, 16 host cores, 400 k-particle sample) | 6.68e+05 /s | @REF@ /s -> e2e speed-up x@SPD@ |
| per step: tree / density / hydro / gravity | 78.9 / 109.6 / 136.4 / 238.3 ms | @BLK@ ms |
| FP64 roofline (measured peak 34.7 TFLOP/s) | 2.85 TFLOP/s = 8.2 % | @FP@ TFLOP/s = @FPF@ % |
| HBM roofline of the dominant pass (gravity, 180 B/particle algorithmic) | 50.7 GB/s = 0.77 % | @HB@ GB/s = @HBF@ % of 6552 (by construction: the pass is FP64 / issue bound); ncu DRAM traffic 1.1 kB per particle -- the interaction lists now travel through HBM on purpose, 380 GB/s while the pass runs |
| 2 GPUs (64 Mi) | value 2.18e+08, e2e 1.5e+08 | value @V2@ /s (@MS2@ ms), e2e @E2@, resident @R2@ |
| 4 GPUs (64 Mi) | value 3.86e+08, e2e 2.18e+08, resident 3.45e+08 | value @V4@ /s (@MS4@ ms), e2e @E4@, resident @R4@ |
| 8 GPUs (64 Mi) | value 6.09e+08 (110 ms), e2e 2.68e+08, resident 4.96e+08 | value @V8@ /s (@MS8@ ms), e2e @E8@ (PCIe bound: every rank still ships its chunk in and out), resident @R8@ |

