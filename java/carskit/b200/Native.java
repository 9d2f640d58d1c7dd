// Native.java -- JNI binding of include/carskit_b200.h, one static native method per C entry point.
//
// Reference side of the drop-in boundary: a recommender subclass overrides only `protected void buildModel()`
// (src/carskit/generic/Recommender.java:1088) and calls these methods; jni/carskit_b200_jni.c is the C half.
// Nothing in this repository's image can compile Java (no JDK); the C half is compiled against jni/stub/jni.h and
// exercised with a fake JNIEnv by tests/test_jni.py.
package carskit.b200;

public final class Native {
    static {
        System.loadLibrary("carskit_b200_jni"); // links libcarskit_b200.so (sm_100a); there is no CPU fallback
    }

    private Native() {
    }

    // enum cars_model
    public static final int PMF = 0, BIASEDMF = 1, CAMF_C = 2, CAMF_CI = 3, CAMF_CU = 4, FM = 5, CAMF_CUCI = 6, CAMF_ICS = 7, CAMF_LCS = 8, CAMF_MCS = 9, SVDPP = 10;
    // enum cars_mode: EXACT is serial-equivalent (P, Q, biases bit-identical to the Java loop); FAST is hogwild
    public static final int EXACT = 0, FAST = 1;
    // enum cars_combine (multi-GPU item-block combine)
    public static final int COMBINE_MEAN = 0, COMBINE_SUM = 1, COMBINE_TOUCHED = 2;

    /** cudaGetDeviceCount; 0 when there is no usable device (cars_create would fail with CARS_E_NO_DEVICE). */
    public static native int deviceCount();

    /** cars_version(). */
    public static native String version();

    /**
     * cars_create. u/j/ctx/r: one entry per training rating in the order of `for (MatrixEntry me : trainMatrix)`
     * (CAMF_CI.java:80); ctx == null for PMF / BiasedMF (iterate the 2-D `train`). ctxPtr/ctxCond: CSR form of what
     * getConditions(ctx) yields (ContextRecommender.java:53-61). The reg* values are the static FLOAT fields widened
     * with a (double) cast by the caller (IterativeRecommender.java:40). gpuIds: null or one id = single GPU; more
     * than one = ONE handle drives all of them (users sharded by range, item block all-reduced inside epoch()).
     * emptyConditions: rateDao.getEmptyContextConditions() for CAMF_ICS / LCS / MCS, null otherwise (its length is also
     * CAMF_MCS's numContextDims). numContextFactors: CAMF_LCS's `-f` (CAMF_LCS.java:38), 0 otherwise.
     * The arrays are copied to the device during the call; nothing stays pinned. Returns the handle.
     */
    public static native long create(int model, int mode, int numUsers, int numItems, int numConditions, int numContexts,
                                     int numFactors, int[] u, int[] j, int[] ctx, double[] r, int[] ctxPtr, int[] ctxCond,
                                     double globalMean, double regU, double regI, double regB, double regC,
                                     int[] gpuIds, int combine, double fastMaxConc, int[] emptyConditions,
                                     int numContextFactors);

    /**
     * cars_upload: flat row-major arrays as initModel() made them; null where the model has no such member. `ccSim` is the
     * model's own extra array: ccMatrix_ICS [C x C], cfMatrix_LCS [C x numF], cVector_MCS [C] or SVD++'s Y [numItems x F].
     */
    public static native void upload(long h, double[] P, double[] Q, double[] userBias, double[] itemBias,
                                     double[] condBias, double[] icBias, double[] ucBias, double[] ccSim);

    /** cars_epoch: one pass over the training ratings; returns the reference's `loss` (already * 0.5); NaN/Inf are returned. */
    public static native double epoch(long h, double lRate);

    /** cars_download: the trained arrays, so that the inherited predict()/evalRatings()/evalRankings() see them. */
    public static native void download(long h, double[] P, double[] Q, double[] userBias, double[] itemBias,
                                       double[] condBias, double[] icBias, double[] ucBias, double[] ccSim);

    /** cars_predict: batched Recommender.predict(u, j, c, bound) (Recommender.java:306-317) on the resident model. */
    public static native void predict(long h, int[] u, int[] j, int[] ctx, boolean bound, double minRate, double maxRate,
                                      double[] out);

    /** cars_eval_ratings: {sum |err|, sum err^2} over a test set (Recommender.java:518-545). */
    public static native double[] evalRatings(long h, int[] u, int[] j, int[] ctx, double[] r, double minRate, double maxRate);

    /**
     * cars_rank_topn: the scoring loop + stable descending sort + subList(0, numRecs) of evalRankings()
     * (Recommender.java:797-824) for a batch of (user, context) queries. cand = candItems in the iteration order of the
     * reference's HashSet (:704); ratedPtr/ratedItems = CSR lists of the items each query's user rated in that context
     * in the training set (:792). outItems/outScores are [numQueries x numRecs]; outCount = rankedItems.size();
     * outKept = itemScores.size() before the cut.
     */
    public static native void rankTopN(long h, int[] qu, int[] qc, int[] cand, long[] ratedPtr, int[] ratedItems,
                                       double binThold, int numRecs, int[] outItems, double[] outScores, int[] outCount,
                                       int[] outKept);

    public static native void destroy(long h);

    // ---- FM (FM.java: ALS, not SGD) -------------------------------------------------------------------------------
    /** cars_fm_create: u/j/ctx/r as above; regLw/regLf = the floats of `FM=-lw .. -lf ..` widened to double. */
    public static native long fmCreate(int numUsers, int numItems, int numConditions, int numContexts, int numFactors,
                                       int numContextDims, int[] u, int[] j, int[] ctx, double[] r, double regLw,
                                       double regLf, int device);

    /** cars_fm_upload (w0 = w0w[0], w = w0w[1..p], V [p x k] row-major) followed by cars_fm_prepare (FM.java:118-146). */
    public static native void fmUploadAndPrepare(long h, double w0, double[] w, double[] V);

    /** cars_fm_iteration: w0 step, w steps, V steps (FM.java:148-219); returns 0.05 * (sum e^2 + regLw terms). */
    public static native double fmIteration(long h);

    /** cars_fm_download: returns w0; w and V are filled. */
    public static native double fmDownload(long h, double[] w, double[] V);

    public static native void fmPredict(long h, int[] u, int[] j, int[] ctx, boolean bound, double minRate, double maxRate,
                                        double[] out);

    public static native void fmDestroy(long h);
}
