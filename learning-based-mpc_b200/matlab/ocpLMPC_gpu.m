function [sysHistory,art_refHistory,true_refHistory]...
          =ocpLMPC_gpu(x,x_wp_init,x_wp,x_wp_ref,u_wp,...
                    N,Ts,iterations,options,opt_var,...
                    Kstabil,Q,R,P,T,Mtheta,LAMBDA,PSI,m,...
                    F_x,h_x,F_u,h_u,F_w_N,h_w_N,...
                    sysHistory,art_refHistory,true_refHistory)
% OCPLMPC_GPU  drop-in for functions/ocpLMPC.m (same argument list and outputs); the fmincon call
% of ocpLMPC.m:20-24 is replaced by lbmpc_mex('solve',...).  A and B are the literals of
% models/nominalModel.m:14-21, which the reference's cost/constraint closures use implicitly.
[A,B] = nominal_AB();
model = struct('A',A,'B',B,'K',Kstabil,'Q',Q,'R',R,'P',P,'T',T,'LAMBDA',LAMBDA,'PSI',PSI, ...
               'F_x',F_x,'h_x',h_x,'F_u',F_u,'h_u',h_u,'F_w_N',F_w_N,'h_w_N',h_w_N);
cfg = struct('form','F','variant','LMPC','N',N,'max_batch',1);
h = lbmpc_mex('create', model, cfg);
cleaner = onCleanup(@() lbmpc_mex('destroy', h));
for k = 1:iterations
    if k > 1
        dx = x - x_wp;
    else
        dx = x_wp_init;                                    % ocpLMPC.m:13-17
    end
    out = lbmpc_mex('solve', h, dx, x_wp_ref, [], opt_var(:));   % replaces fmincon, ocpLMPC.m:24
    if out.status ~= 0
        warning('lbmpc:status', 'step %d: solver status %d', k, out.status);
    end
    opt_var = [out.u_or_c; out.theta];
    theta_opt = reshape(opt_var(end-m+1:end), m, 1);
    c = reshape(opt_var(1:m), m, 1);
    art_ref = Mtheta*theta_opt;
    [x, u] = transitionTrue(x, c, x_wp, u_wp, Kstabil, Ts);      % plant stays in MATLAB (ocpLMPC.m:30)
    his = [x-x_wp; u-u_wp];
    sysHistory = [sysHistory his]; %#ok<*AGROW>
    art_refHistory = [art_refHistory art_ref(1:m)];
    true_refHistory = [true_refHistory x_wp_ref];
end
end

function [A,B] = nominal_AB()
% one nominal step on unit vectors recovers the literals of nominalModel.m without duplicating them
n = 4; A = zeros(n); 
for j = 1:n, e = zeros(n,1); e(j) = 1; A(:,j) = nominalModel(e, 0); end
B = nominalModel(zeros(n,1), 1);
end
