python scripts/bench_eval.py all
python scripts/bench_eval.py gk_ais build/variants/libkabc_cap48.so
python scripts/bench_eval.py gk_ais build/variants/libkabc_cap96.so
