bash tools/final_run.sh r02i
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_r02i.log 2>&1; echo "smoke rc=$?"
