export GSM_LIB_PATH=$PWD/build/libgsm_b200_dbg.so
python tools/tune.py C4 60000 2000 128x3 2>&1 | tail -3
