% BUILD_MEX  compile the gateway against the in-tree shared library (run from this directory).
here = fileparts(mfilename('fullpath'));
mex('-R2018a', fullfile(here,'lbmpc_mex.c'), ['-I' fullfile(here,'..','..','include')], ...
    ['-L' fullfile(here,'..')], '-llbmpc_b200', ['LDFLAGS=$LDFLAGS -Wl,-rpath,' fullfile(here,'..')]);
