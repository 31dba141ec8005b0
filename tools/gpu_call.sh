#!/bin/bash
# the call of the moment: whole tree on one B200 (tests, bench, reference arm, config 5, other configs)
TEST_TIMEOUT=1500 bash tools/gpu_round2.sh ${1:-r04} tests bench c5 configs
