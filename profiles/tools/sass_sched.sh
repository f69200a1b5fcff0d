#!/bin/bash
# Compressed view of a SASS listing's schedule: F = FFMA2, L = LDCU.128 (with its destination), m = FMNMX, a = FADD(2), s = FSEL,
# v = MOV, i = IMAD; one output line per basic block.  Usage: sass_sched.sh file.sass [first_line last_line]
awk -v lo=${2:-1} -v hi=${3:-100000} 'NR>=lo && NR<=hi {split($0,a," "); op=a[1]; if (op ~ /^@/) op=a[2]; printf "%d:%s ", NR, op; if (op ~ /LDCU/) printf "[%s] ", a[2]; if (op ~ /BRA|BSSY|BSYNC|WARPSYNC|EXIT/) printf "\n"}' $1 | sed 's/[0-9]*:FFMA2 /F /g; s/[0-9]*:FMNMX3* /m /g; s/[0-9]*:FADD2* /a /g; s/[0-9]*:FSEL /s /g; s/[0-9]*:MOV /v /g; s/[0-9]*:IMAD[.A-Z0-9]* /i /g; s/[0-9]*:LDCU.128 /L/g'
