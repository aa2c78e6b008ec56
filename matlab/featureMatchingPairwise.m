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
    matches = aps_featureMatching_mex('pairwise', allDescriptors, numImg, input.Matchingthreshold, input.Ratiothreshold);
end
