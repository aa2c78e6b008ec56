function [allMatches, numMatches, tforms] = imageMatching(input, n, keypoints, matchesAll, imagesProcessed) %#ok<INUSD>
    %IMAGEMATCHING  Drop-in replacement of PP/imageMatching/imageMatching.m (custom 'ransac' branch).
    %   Same signature and outputs: top-m partner selection, RANSAC homography per candidate pair
    %   ('projective'), acceptance ni > 8 + 0.3*nf, tforms{i,j} / tforms{j,i} = model / inv(model).
    %   All candidate pairs and all trials run in ONE GPU call (aps_imageMatching_mex -> libapsmatch.so);
    %   no CPU fallback.  The showKeypointsPlot visualisation of refineMatch is not reproduced.
    arguments
        input (1, 1) struct
        n (1, 1) double {mustBeInteger, mustBePositive}
        keypoints cell
        matchesAll cell
        imagesProcessed cell
    end
    if input.useMATLABImageMatching == 1 || ~strcmpi(input.imageMatchingMethod, 'ransac') || ...
            ~strcmpi(input.transformationType, 'projective')
        error('apsmatch:args', 'GPU path covers imageMatchingMethod = ''ransac'', transformationType = ''projective''.');
    end
    [allMatches, numMatches, tforms] = aps_imageMatching_mex(keypoints, matchesAll, input.mBrownLowe, ...
        input.maxDistance, input.inliersConfidence, input.maxIter);
end
