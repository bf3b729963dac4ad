"""The reference's known-answer tests, restated as data.

frcfrc/unifrac_test.go:12-31  TestUniFrac_simple    -> [6/9]
frcfrc/unifrac_test.go:33-53  TestUniFrac_complex   -> [19/28, 16/22, 1]
frcfrc/unifrac_test.go:55-74  TestUniFrac_weighted  -> [22/36]
testdata/run.sh:3-16          six CLI runs diffed byte-for-byte against *.want
"""
UNIT = [
    dict(name="simple", tree="(s2:3,s1:1,s3:5);", weighted=False,
         abnd=[{"s1": 1, "s2": 1}, {"s3": 1, "s2": 1}], want=[6.0 / 9.0]),
    dict(name="complex", tree="((s1:1,s2:3,s3:5):3,(s4:2,s5:2,s6:2):4,(s7:3,s8:2,s9:1):5);", weighted=False,
         abnd=[{"s1": 1, "s2": 1, "s5": 1, "s9": 1}, {"s3": 1, "s4": 1, "s5": 1, "s6": 1}, {"s7": 1, "s9": 1}],
         want=[19.0 / 28.0, 16.0 / 22.0, 1.0]),
    dict(name="weighted", tree="((s1:1,s2:3):2,(s3:2,s4:5):1);", weighted=True,
         abnd=[{"s1": 4, "s2": 1}, {"s3": 3, "s2": 2}], want=[22.0 / 36.0]),
]

# (fixture, sparse, weighted): testdata/run.sh
CLI = [("uwtd1", False, False), ("uwtd1", True, False), ("uwtd2", False, False), ("uwtd2", True, False),
       ("wtd", False, True), ("wtd", True, True)]


def sparse_text(abnd):
    return "".join("\t".join(f"{k}:{v}" for k, v in m.items()) + "\n" for m in abnd)


# (fixture, table suffix, sparse, want tag, weighted, normalize): tests/golden/make_golden.py (oracle answers on
# seeded synthetic inputs, committed).  normalize: 1 = default, 0 = flag -l as coded in the reference (what the CLI
# prints for -l), 2 = flag -l as documented (FRCFRC_L=documented)
SYNTH = [("synth_a", ".sparse", True, "uw", False, 1), ("synth_a", ".dense", False, "uw", False, 1),
         ("synth_a", ".sparse", True, "w", True, 1), ("synth_a", ".dense", False, "w", True, 1),
         ("synth_a", ".sparse", True, "wl", True, 0), ("synth_a", ".sparse", True, "wl2", True, 2),
         ("synth_b", ".sparse", True, "uw", False, 1), ("synth_b", ".sparse", True, "w", True, 1),
         ("synth_b", ".sparse", True, "wl", True, 0), ("synth_b", ".sparse", True, "wl2", True, 2)]
