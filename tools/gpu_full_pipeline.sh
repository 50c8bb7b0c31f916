#!/bin/bash
# evidence: the reference's whole pipeline, all-reference vs the three GPU entry points (oracle/_ref/poppy_dropin full)
mkdir -p gpurun_out
export LD_LIBRARY_PATH=$PWD/poppy_b200:$(python -c "import glob,sys; print(':'.join(sum([glob.glob(p+'/nvidia/cuda_runtime/lib') for p in sys.path],[])))"):/usr/local/cuda/lib64:$LD_LIBRARY_PATH
oracle/_ref/poppy_dropin full oracle/_ref/full/c1 60 2>/dev/null | grep '^{' > gpurun_out/full_pipeline_c1.json; cat gpurun_out/full_pipeline_c1.json
oracle/_ref/poppy_dropin full oracle/_ref/full/c3 120 2>/dev/null | grep '^{' > gpurun_out/full_pipeline_c3.json; cat gpurun_out/full_pipeline_c3.json
