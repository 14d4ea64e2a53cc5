for l in ${LEVELS:-1 2 3 4 5 6}; do echo level $l; VD_DEBUG_SKIP_EPILOGUE=$l python scripts/head_stage_time.py ${WL:-voc416_b64}; done
python scripts/head_stage_time.py ${WL:-voc416_b64}
