function matches = featureMatchingPairwise(input, allDescriptors, numImg)
    %FEATUREMATCHINGPAIRWISE  Drop-in replacement of PP/featureMatching/featureMatchingPairwise.m.
    %   Runs getMatches' matchFeaturesScratch 'Exhaustive' branch (Unique = true) for every image pair
    %   i<j on the GPU in one call.  (The MathWorks matchFeatures branch, input.useMATLABFeatureMatch = 1,
    %   is closed source; with this file on the path the exhaustive scratch semantics are used.)
    arguments
        input struct
        allDescriptors cell
        numImg (1, 1) {mustBeNumeric, mustBeFinite, mustBePositive}
    end
    % PP/inputs.m:47-49 defaults select MathWorks matchFeatures (closed source) and 'Approximate': say so instead of
    % changing the match lists silently; input.apsAcceptScratchSemantics = 1 acknowledges and silences this.
    if ~(isfield(input, 'apsAcceptScratchSemantics') && input.apsAcceptScratchSemantics)
        if isfield(input, 'useMATLABFeatureMatch') && input.useMATLABFeatureMatch == 1
            warning('apsmatch:semantics', ['input.useMATLABFeatureMatch = 1 selects MathWorks matchFeatures in the ' ...
                'reference; this GPU path runs the matchFeaturesScratch semantics (featureMatchingPairwise.m:108-117).']);
        end
        if isfield(input, 'Matchingmethod') && strcmpi(input.Matchingmethod, 'Approximate')
            warning('apsmatch:semantics', ['Matchingmethod = ''Approximate'': float descriptors are matched by the ' ...
                'exact search (matchFeaturesScratch.m:128-163 approximates it); binary descriptors run the ' ...
                'exhaustive search as in the reference (:611).']);
        end
    end
    matches = aps_featureMatching_mex('pairwise', allDescriptors, numImg, input.Matchingthreshold, input.Ratiothreshold);
end
