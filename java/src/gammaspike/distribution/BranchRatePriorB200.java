package gammaspike.distribution;

import beast.base.core.Description;
import gammaspike.b200.GsmB200;
import gammaspike.b200.GsmContext;

/** BranchRatePrior (== the template's core Prior + LogNormal(meanInRealSpace, M = 1, S = GSMclockSD),
 *  fxtemplates/GammaSpikeClockModel.xml:60-62) evaluated by libgsm_b200 (UNCOMPILED here). */
@Description("Log normal prior on the branch rates; evaluated by libgsm_b200")
public class BranchRatePriorB200 extends BranchRatePrior {
    private GsmContext ctx;

    @Override
    public void initAndValidate() {
        super.initAndValidate();
        ctx = GsmContext.of(treeInput.get());
        ctx.ratePrior = this;
    }

    /** BRP:29-65 */
    @Override
    public double calculateLogP() {
        logP = ctx.eval()[GsmB200.O_RATE_LOGP];
        return logP;
    }

    @Override
    protected boolean requiresRecalculation() {
        ctx.invalidate();
        return true;
    }

    @Override
    public void restore() {
        ctx.invalidate();
        super.restore();
    }
}
