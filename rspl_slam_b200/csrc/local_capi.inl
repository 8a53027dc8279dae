extern "C" int rspl_ba_local_batch_upload(RsplBaContext* c, const RsplLocalBatch*) { return fail(c, RSPL_BA_ERR_UNSUPPORTED, "local BA not built yet"); }
extern "C" int rspl_ba_local_batch_solve(RsplBaContext* c, const RsplBaOptions*) { return fail(c, RSPL_BA_ERR_UNSUPPORTED, "local BA not built yet"); }
extern "C" int rspl_ba_local_batch_download(RsplBaContext* c, RsplLocalBatchResult*) { return fail(c, RSPL_BA_ERR_UNSUPPORTED, "local BA not built yet"); }
extern "C" int rspl_ba_local_batch(RsplBaContext* c, const RsplLocalBatch*, const RsplBaOptions*, RsplLocalBatchResult*) { return fail(c, RSPL_BA_ERR_UNSUPPORTED, "local BA not built yet"); }
