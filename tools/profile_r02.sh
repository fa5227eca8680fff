#!/bin/bash
# Round-2 ncu evidence (run under gpurun, 1 GPU).  Eager network evaluations of the deep NCSN++ at batch 256 + short samplers:
#   r02_launches.csv       every launch of the 2nd evaluation with its duration (cold cache, serialised: compare SHARES)
#   r02_gemm_traffic.csv   DRAM bytes + duration of every conv_gemm_umma launch of the 2nd evaluation
#   r02_*.ncu-rep          --set full captures: GroupNorm-epilogue GEMMs, plain GEMMs, and every HBM-bound kernel family
O=gpurun_out
mkdir -p $O
# ONLY=name1,name2: re-capture just those --set full reports (skips the launch list and the traffic pass)
if [ -z "$ONLY" ]; then
N=$(python tools/prof_forward.py 256 | awk '/launches/{print $2}')
PER=$((N / 2))
SKIP=$((PER + 3))
echo "launches per forward: $PER" | tee $O/r02_prof.log
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c $PER --csv --log-file $O/r02_launches.csv \
    python tools/prof_forward.py 256 >> $O/r02_prof.log 2>&1
NG=$(grep -c conv_gemm_umma $O/r02_launches.csv)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:conv_gemm_umma \
    -s $NG -c $NG --csv --log-file $O/r02_gemm_traffic.csv python tools/prof_forward.py 256 >> $O/r02_prof.log 2>&1
fi
# (gpurun only copies back 64 MiB: the reports stay on the box, their raw pages -- every metric of the capture, one CSV
# row per launch -- come back; KEEP_REP=name1,name2 also brings those .ncu-rep files home for the source view)
T=/tmp/r02_prof
mkdir -p $T
full() {  # name, kernel regex (mangled), skip, count, script...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  if [ -n "$ONLY" ]; then case ",$ONLY," in *",$name,"*) ;; *) return ;; esac; fi
  ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:$rx -s $skip -c $cnt -f -o $T/r02_$name "$@" >> $O/r02_prof.log 2>&1
  ncu -i $T/r02_$name.ncu-rep --page raw --csv > $O/r02_${name}_raw.csv 2>> $O/r02_prof.log
  case ",$KEEP_REP," in *",$name,"*) cp $T/r02_$name.ncu-rep $O/ ;; esac
}
full gemm_gnf_16x16_pair   'conv_gemm_umma_kernelILi256ELi2ELi1ELi2' 10 1 python tools/prof_forward.py 256
full gemm_gnf_32x32_pairs4 'conv_gemm_umma_kernelILi128ELi2ELi2ELi2ELb1' 4 1 python tools/prof_forward.py 256
full gemm_gnf_32x32_k1152  'conv_gemm_umma_kernelILi128ELi2ELi2ELi2ELb1' 0 1 python tools/prof_forward.py 256
full gemm_gnf_8x8_dual     'conv_gemm_umma_kernelILi256ELi2ELi1ELi1' 6 1 python tools/prof_forward.py 256
full gemm_lin_32x32_halo   'conv_gemm_umma_kernelILi128ELi0ELi2ELi2ELb1' 8 1 python tools/prof_forward.py 256
full gemm_lin_16x16_pair   'conv_gemm_umma_kernelILi256ELi0ELi1ELi2' 8 1 python tools/prof_forward.py 256
full gn_apply              'gn_apply_kernelILi0E' 30 1 python tools/prof_forward.py 256
full gn_fir_down           'gn_fir_tiled_kernelILi1E' 0 1 python tools/prof_forward.py 256
full gn_fir_up             'gn_fir_tiled_kernelILi2E' 2 1 python tools/prof_forward.py 256
full gn_coef               'gn_coef_kernel' 30 1 python tools/prof_forward.py 256
full gn_small              'gn_small_kernel' 20 1 python tools/prof_forward.py 256
full gn_qkv                'gn_qkv_kernel' 4 1 python tools/prof_forward.py 256
full attn                  'attn256_kernel' 4 1 python tools/prof_forward.py 256
full im2col_fir_down       'im2col_fir_down' 1 1 python tools/prof_forward.py 256
full cld_step              'cld_step_c3_kernel' 3 1 python tools/prof_sampler.py cld 256 4
full blur_step             'blur_step_kernel' 1 1 python tools/prof_sampler.py blur 256 4
full dct32                 'dct32_kernel' 0 1 python tools/prof_sampler.py blur 256 4
ls -la $T/*.ncu-rep $O/r02_* | awk '{print $5, $9}' | tee -a $O/r02_prof.log
