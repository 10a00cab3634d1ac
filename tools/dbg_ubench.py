import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import rna_sequence_diff_patch_b200 as R
eng = R.Engine(0)
names = {0: "IADD3", 1: "VIADDMNMX", 2: "VIADDMNMX.S16x2", 3: "PRMT", 4: "DADD", 5: "IMAD", 6: "VIMNMX3", 8: "VIMNMX.S16x2", 9: "IMAD.HI.U32", 10: "IMAD.WIDE.U32 (+LOP)", 11: "SHF.L.W"}
for k, n in names.items():
    print(f"{n:24s} {eng.ubench(k) * 1e-12:7.2f} Tops/s", flush=True)
