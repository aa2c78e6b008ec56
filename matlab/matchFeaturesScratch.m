function [matches, matchMetric] = matchFeaturesScratch(F1, F2, varargin)
    %MATCHFEATURESSCRATCH  Drop-in replacement of PP/featureMatching/matchFeaturesScratch.m (GPU, exhaustive).
    %   [matches, matchMetric] = matchFeaturesScratch(F1, F2, 'Method', 'Exhaustive' | 'Approximate', ...
    %       'MatchThreshold', T, 'MaxRatio', r, 'Unique', tf, <the reference's Approx* options>)
    %   Same inputs and outputs as the reference: F1/F2 float [N x D] (single/double), binaryFeatures objects, or
    %   logical / 0-1 uint8 bit matrices [N x Dbits]; matches [K x 2] uint32 rows of F1 / F2, matchMetric [K x 1]
    %   (SSD, or percent of mismatched bits).  Nearest-2 search, the "normalise iff max|.| > 2" rule, ratio and
    %   threshold tests and the greedy Unique pass (reference lines 105-126, 169-215) run on the GPU in
    %   aps_matchFeatures_mex; this file only parses the options and decides the descriptor kind (lines 237-292).
    %   'Approximate': binary descriptors run the exhaustive search (the reference's own branch is the exhaustive OMP MEX,
    %   line 611; the LSH options are accepted and ignored as there); float descriptors follow 'ApproxFloatNNMethod':
    %   'pca2nn' (the parser default, PCA-48 + cosine, lines 130-141), 'subsetpdist2' and 'kdtree' (Euclidean searches, 142-155).
    opt = struct('Method', 'Exhaustive', 'MatchThreshold', 3.5, 'MaxRatio', 0.6, 'Unique', true, 'ApproxFloatNNMethod', 'pca2nn');
    accepted = {'ApproxNumTables', 'ApproxBitsPerKey', 'ApproxProbes', 'ApproxKDBucketSize', ...
                'ApproxKDTreeLeafSize', 'Approx.NumTables', 'Approx.BitsPerKey', 'Approx.Probes'};
    if mod(numel(varargin), 2) ~= 0
        error('apsmatch:args', 'Options must be name-value pairs.');
    end
    for a = 1:2:numel(varargin)
        name = char(varargin{a});
        hit = find(strcmpi(name, fieldnames(opt)), 1);
        if ~isempty(hit)
            names = fieldnames(opt);
            opt.(names{hit}) = varargin{a + 1};
        elseif ~any(strcmpi(name, accepted))
            error('apsmatch:args', 'Unknown option ''%s''.', name);
        end
    end
    method = lower(char(opt.Method));
    if ~any(strcmp(method, {'exhaustive', 'approximate'}))
        error('Unknown Method: %s', char(opt.Method));
    end
    if ~(isnumeric(opt.MatchThreshold) && isscalar(opt.MatchThreshold) && opt.MatchThreshold >= 0)
        error('apsmatch:args', 'MatchThreshold must be a non-negative scalar.');
    end
    if ~(isnumeric(opt.MaxRatio) && isscalar(opt.MaxRatio) && opt.MaxRatio > 0 && opt.MaxRatio <= 1)
        error('apsmatch:args', 'MaxRatio must be a scalar in (0, 1].');
    end
    if ~islogical(opt.Unique)
        error('apsmatch:args', 'Unique must be logical.');
    end

    if isa(F1, 'binaryFeatures') && isa(F2, 'binaryFeatures')
        kind = 1; A = F1.Features; B = F2.Features;                  % packed bytes, nBits = 8 * size(A, 2)
    elseif isBits(F1) && isBits(F2)
        kind = 2; A = uint8(F1); B = uint8(F2);                      % unpacked bits, packed MSB-first on the device
    else
        kind = 0; A = F1; B = F2;
        if ~isa(A, 'single') && ~isa(A, 'double'); A = single(A); end
        if ~isa(B, 'single') && ~isa(B, 'double'); B = single(B); end
    end
    nnMethod = 0;                                        % aps_method: 0 exhaustive, 1 subsetpdist2, 2 kdtree
    if kind == 0 && strcmp(method, 'approximate')
        switch lower(char(opt.ApproxFloatNNMethod))
            case 'subsetpdist2'
                nnMethod = 1;
            case 'kdtree'
                nnMethod = 2;
            case 'pca2nn'
                nnMethod = 3;
            otherwise
                error('Select a approximate method');
        end
    end
    [matches, matchMetric] = aps_matchFeatures_mex(A, B, kind, double(opt.MatchThreshold), double(opt.MaxRatio), ...
                                                   logical(opt.Unique), nnMethod);
end

function tf = isBits(F)
    tf = islogical(F) || (isa(F, 'uint8') && all(F(:) == 0 | F(:) == 1));
end
